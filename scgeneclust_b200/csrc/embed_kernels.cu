// Gene-by-cell embedding kernels: ingest, exact count sums, analytic Pearson residuals, and the randomized-SVD
// passes  Y = M Q  /  Y = M^T Q  over the implicit residual matrix
//     M[c,g] = clip((x - mu)/sqrt(mu + mu^2/theta)) - row_shift[c] - col_shift[g],   mu = r_c * s_g / T
// (scanpy normalize_pearson_residuals + sklearn PCA centring + randomized_svd BLAS passes, reached from
// scGeneClust/pp/_preprocessing.py:29,52,54).  The residual matrix is never materialised for the fast path:
// the uint16 counts are re-read from HBM once per pass and the residual is produced in registers.
//
// This file holds the SIMT (FFMA) implementation of the passes; rsvd_tc.cu holds the tcgen05 one.
#include "common.cuh"

namespace gc {

// ---------------------------------------------------------------------------------------------------------
// ingest
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t to_count(float v, int32_t *flags) {
    if (!(v >= 0.f) || v != floorf(v)) { flags[0] = 1; return 0; }   // _validation.py:69  X % 1 == 0
    if (v > 65535.f) { flags[1] = 1; return 65535; }
    return (uint16_t)v;
}

// flags: [0] value not a non-negative integer, [1] value (or sum of duplicates) above 65535, [2] column index outside
// [0, G), [3] row not canonical (indices not strictly increasing: unsorted or duplicate entries).
// ACCUMULATE = false: plain stores, valid for canonical rows only - a set flags[3] tells the caller to redo the block
// with ACCUMULATE = true, where duplicates are summed like scipy's toarray() (scGeneClust/_model.py:124) by a 32-bit
// atomic add on the word that holds the uint16.
template <bool ACCUMULATE>
__global__ void csr_to_counts_kernel(const float *__restrict__ data, const int32_t *__restrict__ indices,
                                     const int64_t *__restrict__ indptr, int64_t n_rows, int64_t row0, int64_t G,
                                     uint16_t *__restrict__ counts, int64_t ld, int32_t *flags) {
    // one warp per CSR row
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    const int64_t base = indptr[0];
    const int64_t e0 = indptr[r] - base, e1 = indptr[r + 1] - base;
    uint16_t *dst = counts + (row0 + r) * ld;
    for (int64_t e = e0 + lane; e < e1; e += 32) {
        const int32_t col = indices[e];
        if (col < 0 || col >= G) { flags[2] = 1; continue; }
        const uint16_t v = to_count(data[e], flags);
        if (!ACCUMULATE) {
            if (e > e0 && indices[e - 1] >= col) flags[3] = 1;
            dst[col] = v;
        } else if (v) {
            const int64_t pos = (row0 + r) * ld + col;                       // counts is 4-byte aligned, ld is even
            const unsigned shift = (unsigned)(pos & 1) * 16u;
            const unsigned old = atomicAdd(reinterpret_cast<unsigned int *>(counts) + (pos >> 1), (unsigned)v << shift);
            if (((old >> shift) & 0xffffu) + v > 65535u) flags[1] = 1;
        }
    }
}

__global__ void dense_to_counts_kernel(const float *__restrict__ src, int64_t ld_src, int64_t n_rows, int64_t G,
                                       int64_t row0, uint16_t *__restrict__ counts, int64_t ld, int32_t *flags) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (g < G && r < n_rows) counts[(row0 + r) * ld + g] = to_count(src[r * ld_src + g], flags);
}

// ---------------------------------------------------------------------------------------------------------
// exact integer sums.  Tile = 128 rows x 2048 columns per CTA, thread = 8 adjacent columns (one 16-byte load
// per row).  Integer atomics => deterministic.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSumRows = 128;
__global__ void __launch_bounds__(256)
count_sums_kernel(const uint16_t *__restrict__ counts, int64_t C, int64_t G, int64_t ld,
                  unsigned long long *__restrict__ row_sum, unsigned long long *__restrict__ col_sum,
                  uint32_t *__restrict__ col_max) {
    __shared__ unsigned int srow[kSumRows];
    const int64_t r0 = (int64_t)blockIdx.y * kSumRows;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;      // blockDim.x = 64 .. 256, see the launch
    for (int r = threadIdx.x; r < kSumRows; r += blockDim.x) srow[r] = 0;
    __syncthreads();
    unsigned int cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned int mx01 = 0, mx23 = 0, mx45 = 0, mx67 = 0;       // per-gene maxima, two uint16 lanes per word (__vmaxu2)
    const bool active = g0 < ld;  // ld is a multiple of 8, padded columns hold zeros
    const int rows = (int)min((int64_t)kSumRows, C - r0);
    // four rows per iteration (four 16-byte loads in flight per thread); the row sums of a warp are folded by the
    // hardware warp reduction (REDUX) instead of five shuffle steps per row
    for (int rb = 0; rb < rows; rb += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u] = make_uint4(0u, 0u, 0u, 0u);
            if (active && rb + u < rows) v[u] = *reinterpret_cast<const uint4 *>(counts + (r0 + rb + u) * ld + g0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned int w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            mx01 = __vmaxu2(mx01, v[u].x); mx23 = __vmaxu2(mx23, v[u].y);
            mx45 = __vmaxu2(mx45, v[u].z); mx67 = __vmaxu2(mx67, v[u].w);
            unsigned int rs = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned int lo = w[k] & 0xffffu, hi = w[k] >> 16;
                cs[2 * k] += lo; cs[2 * k + 1] += hi; rs += lo + hi;
            }
            rs = __reduce_add_sync(0xffffffffu, rs);
            if ((threadIdx.x & 31) == 0 && rs) atomicAdd(&srow[rb + u], rs);
        }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < rows; r += blockDim.x)
        if (srow[r]) atomicAdd(&row_sum[r0 + r], (unsigned long long)srow[r]);
    if (active) {
        const unsigned int mx[8] = {mx01 & 0xffffu, mx01 >> 16, mx23 & 0xffffu, mx23 >> 16,
                                    mx45 & 0xffffu, mx45 >> 16, mx67 & 0xffffu, mx67 >> 16};
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (g0 + k < G && cs[k]) {
                atomicAdd(&col_sum[g0 + k], (unsigned long long)cs[k]);
                if (col_max) atomicMax(&col_max[g0 + k], mx[k]);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------
// materialised residuals (ps path: adata.X = residuals) - elementwise, 8 columns per thread
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pearson_residuals_kernel(const uint16_t *__restrict__ counts, int64_t C, int64_t G, int64_t ld,
                         const float *__restrict__ row_sum_f, const float *__restrict__ col_sum_f, float inv_total,
                         float inv_theta, float clip, float *__restrict__ Z, int64_t ldz) {
    const int64_t g0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
    const int64_t c = blockIdx.y;
    if (g0 >= G) return;
    const float r = row_sum_f[c];
    const uint4 v = *reinterpret_cast<const uint4 *>(counts + c * ld + g0);
    const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (g0 + k < G) {
            const float x = (float)((w[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
            Z[c * ldz + g0 + k] = pearson_z(x, r, col_sum_f[g0 + k] * inv_total, inv_theta, clip);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// residual sums in float64, deterministic:
//   rows: one warp per cell, fixed lane-strided order + fixed shuffle tree
//   cols: thread per gene over a slab of cells, slab partials reduced in slab order by a second kernel
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
residual_row_sums_kernel(const uint16_t *__restrict__ counts, int64_t C, int64_t G, int64_t ld,
                         const float *__restrict__ row_sum_f, const float *__restrict__ col_sum_f, float inv_total,
                         float inv_theta, float clip, double *__restrict__ zrow_sum) {
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    const float r = row_sum_f[c];
    double acc = 0.0;
    for (int64_t g0 = (int64_t)lane * 8; g0 < G; g0 += 256) {
        const uint4 v = *reinterpret_cast<const uint4 *>(counts + c * ld + g0);
        const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (g0 + k < G) {
                const float x = (float)((w[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
                acc += (double)pearson_z(x, r, col_sum_f[g0 + k] * inv_total, inv_theta, clip);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) zrow_sum[c] += acc;
}

constexpr int kColSlabs = 32;
__global__ void __launch_bounds__(256)
residual_col_partials_kernel(const uint16_t *__restrict__ counts, int64_t C, int64_t G, int64_t ld,
                             const float *__restrict__ row_sum_f, const float *__restrict__ col_sum_f,
                             float inv_total, float inv_theta, float clip, double *__restrict__ partial) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int slab = blockIdx.y;
    const int64_t rows_per = ceil_div<int64_t>(C, kColSlabs);
    const int64_t c_lo = slab * rows_per, c_hi = min(C, c_lo + rows_per);
    if (g >= G) return;
    const float sT = col_sum_f[g] * inv_total;
    double acc = 0.0;
    for (int64_t c = c_lo; c < c_hi; ++c)
        acc += (double)pearson_z((float)counts[c * ld + g], row_sum_f[c], sT, inv_theta, clip);
    partial[(int64_t)slab * G + g] = acc;
}

__global__ void residual_col_reduce_kernel(const double *__restrict__ partial, int64_t G,
                                           double *__restrict__ zcol_sum) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= G) return;
    double acc = 0.0;
    for (int s = 0; s < kColSlabs; ++s) acc += partial[(int64_t)s * G + g];
    zcol_sum[g] += acc;
}

// ---------------------------------------------------------------------------------------------------------
// randomized-SVD pass, SIMT version.
//   trans = 0: out rows = cells (tile of 64), contraction over genes in steps of 32
//   trans = 1: out rows = genes (tile of 64), contraction over cells in steps of 32
// 256 threads, each a 4x4 register tile of the 64 x 64 output tile; A tile (64 rows x 32 k, stored k-major) and
// Q tile (32 k x 64) double-buffered in shared memory.  The contraction range is split in `n_split` slabs;
// CTA (tile, split) writes its partial tile to `partials[split]`, a second kernel adds the slabs in order.
// ---------------------------------------------------------------------------------------------------------
constexpr int kTileM = 64, kTileK = 32, kQ = 64;

struct ApplyArgs {
    const uint16_t *counts; int64_t C, G, ld;
    const float *row_sum_f, *col_sum_f; float inv_total, inv_theta, clip;
    const float *row_shift, *col_shift;
    const float *Qin; float *partials;
    int64_t n_out, n_k; int n_split; int64_t k_per_split;   // k_per_split is a multiple of kTileK
};

template <int TRANS>
__global__ void __launch_bounds__(256)
rsvd_apply_simt_kernel(const ApplyArgs a) {
    __shared__ __align__(16) float sA[2][kTileK][kTileM];
    __shared__ __align__(16) float sQ[2][kTileK][kQ];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * kTileM;
    const int split = blockIdx.y;
    const int64_t k_lo = split * a.k_per_split, k_hi = min(a.n_k, k_lo + a.k_per_split);
    const int n_it = (int)ceil_div<int64_t>(max((int64_t)0, k_hi - k_lo), kTileK);

    // --- producer mapping: each thread converts 8 adjacent genes of one cell per k-step -----------------
    // TRANS=0: tile = 64 cells x 32 genes ; thread -> (cell = tid/4, gene segment = tid%4)
    // TRANS=1: tile = 32 cells x 64 genes ; thread -> (cell = tid/8, gene segment = tid%8)
    const int p_cell = TRANS == 0 ? tid >> 2 : tid >> 3;
    const int p_seg = TRANS == 0 ? tid & 3 : tid & 7;
    float r_c = 0.f, sh_c = 0.f;           // TRANS=0: fixed cell per thread
    float sT[8], sh_g[8];                  // TRANS=1: fixed genes per thread
    if (TRANS == 0) {
        const int64_t c = m0 + p_cell;
        if (c < a.C) { r_c = a.row_sum_f[c]; sh_c = a.row_shift ? a.row_shift[c] : 0.f; }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int64_t g = m0 + p_seg * 8 + e;
            sT[e] = g < a.G ? a.col_sum_f[g] * a.inv_total : 0.f;
            sh_g[e] = (g < a.G && a.col_shift) ? a.col_shift[g] : 0.f;
        }
    }

    uint4 raw = make_uint4(0, 0, 0, 0);
    auto load_raw = [&](int it) {
        const int64_t kk = k_lo + (int64_t)it * kTileK;
        int64_t c, g0;
        if (TRANS == 0) { c = m0 + p_cell; g0 = kk + p_seg * 8; }
        else { c = kk + p_cell; g0 = m0 + p_seg * 8; }
        const bool ok = c < a.C && g0 < a.ld && (TRANS == 0 ? g0 < k_hi : c < k_hi);
        raw = ok ? *reinterpret_cast<const uint4 *>(a.counts + c * a.ld + g0) : make_uint4(0, 0, 0, 0);
    };
    auto store_tile = [&](int it, int buf) {
        const int64_t kk = k_lo + (int64_t)it * kTileK;
        const unsigned int w[4] = {raw.x, raw.y, raw.z, raw.w};
        float z[8];
        if (TRANS == 0) {
            const int64_t c = m0 + p_cell;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int64_t g = kk + p_seg * 8 + e;
                const bool ok = c < a.C && g < k_hi;
                const float x = (float)((w[e >> 1] >> ((e & 1) * 16)) & 0xffffu);
                float v = 0.f;
                if (ok) {
                    v = pearson_z(x, r_c, a.col_sum_f[g] * a.inv_total, a.inv_theta, a.clip) - sh_c;
                    if (a.col_shift) v -= a.col_shift[g];
                }
                z[e] = v;
            }
            // k-major store with a 16-byte-granular XOR swizzle: conflict-free scalar stores, float4 loads
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = p_seg * 8 + e;
                sA[buf][k][p_cell ^ (((k >> 3) & 3) << 3)] = z[e];
            }
        } else {
            const int64_t c = kk + p_cell;
            const bool okc = c < k_hi;   // k_hi <= C
            const float r = okc ? a.row_sum_f[c] : 0.f;
            const float shc = (okc && a.row_shift) ? a.row_shift[c] : 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int64_t g = m0 + p_seg * 8 + e;
                const float x = (float)((w[e >> 1] >> ((e & 1) * 16)) & 0xffffu);
                z[e] = (okc && g < a.G) ? pearson_z(x, r, sT[e], a.inv_theta, a.clip) - shc - sh_g[e] : 0.f;
            }
            float4 *dst = reinterpret_cast<float4 *>(&sA[buf][p_cell][p_seg * 8]);
            dst[0] = make_float4(z[0], z[1], z[2], z[3]);
            dst[1] = make_float4(z[4], z[5], z[6], z[7]);
        }
        // Q tile: rows kk..kk+31 of Qin (ld 64): 2048 floats = 2 float4 per thread
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int idx = tid + h * 256;           // float4 index in the 32 x 16 tile
            const int k = idx >> 4, c4 = idx & 15;
            const int64_t kr = kk + k;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kr < k_hi) q = *reinterpret_cast<const float4 *>(a.Qin + kr * kQ + c4 * 4);
            *reinterpret_cast<float4 *>(&sQ[buf][k][c4 * 4]) = q;
        }
    };

    const int ty = tid >> 4, tx = tid & 15;   // rows ty*4.., cols tx*4..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    if (n_it > 0) {
        load_raw(0);
        store_tile(0, 0);
        __syncthreads();
        for (int it = 0; it < n_it; ++it) {
            const int buf = it & 1;
            if (it + 1 < n_it) load_raw(it + 1);
#pragma unroll
            for (int k = 0; k < kTileK; ++k) {
                const int col = TRANS == 0 ? ((ty * 4) ^ (((k >> 3) & 3) << 3)) : ty * 4;
                const float4 av = *reinterpret_cast<const float4 *>(&sA[buf][k][col]);
                const float4 qv = *reinterpret_cast<const float4 *>(&sQ[buf][k][tx * 4]);
                const float am[4] = {av.x, av.y, av.z, av.w}, qm[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], qm[j], acc[i][j]);
            }
            if (it + 1 < n_it) store_tile(it + 1, buf ^ 1);
            __syncthreads();
        }
    }
    float *out = a.partials + (int64_t)split * a.n_out * kQ;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = m0 + ty * 4 + i;
        if (row < a.n_out)
            *reinterpret_cast<float4 *>(out + row * kQ + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

__global__ void rsvd_reduce_partials_kernel(const float *__restrict__ partials, int64_t n_elem4, int n_split,
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

// split heuristic shared by scratch sizing and launch: fill a multiple of (148 SMs x 4 resident CTAs)
__host__ inline int choose_split(int64_t n_tiles, int64_t n_k) {
    const int64_t slots = (int64_t)kNumSMs * 4;
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 32; ++s) {
        const int64_t k_per = ceil_div<int64_t>(ceil_div<int64_t>(n_k, s), kTileK) * kTileK;
        if (s > 1 && k_per < 8 * kTileK) break;
        const int64_t ctas = n_tiles * s;
        const double eff = (double)ctas / (double)(ceil_div<int64_t>(ctas, slots) * slots);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    return best;
}

// ---------------------------------------------------------------------------------------------------------
// tall-skinny panel algebra (rows x 64 float32 panels)
// ---------------------------------------------------------------------------------------------------------
// Gram P^T P in float64.  The 64 x 64 result is a 16 x 16 grid of 4 x 4 blocks; only the 136 blocks on or above the
// diagonal are computed (one thread each, 160-thread CTAs), the reduce kernel mirrors them.  The row range is split
// over up to 296 CTAs (two per SM): round 1's fixed 512 rows per CTA left a third of the SMs idle on a 50 000-row panel
// and made each CTA's float64 FMA chain 512 rows long (47 us per Gram, 17 per selection).
constexpr int kGramMaxCtas = 296;
constexpr int kGramThreads = 160;
constexpr int kGramBlocks = 136;

__host__ __device__ inline int64_t gram_rows_per_cta(int64_t rows) {
    const int64_t per = (rows + kGramMaxCtas - 1) / kGramMaxCtas;
    return ((per + 31) / 32) * 32;
}

__global__ void __launch_bounds__(kGramThreads)
panel_gram_partial_kernel(const float *__restrict__ P, int64_t rows, int64_t rows_per_cta, double *__restrict__ partial) {
    __shared__ __align__(16) float sP[32][kQ];
    const int t = threadIdx.x;
    const bool active = t < kGramBlocks;
    // block (ty, tx), ty <= tx, row-major over the upper triangle
    int ty = 0, rem = t;
    while (active && rem >= 16 - ty) { rem -= 16 - ty; ++ty; }
    const int tx = active ? ty + rem : 0;
    const int64_t r_lo = (int64_t)blockIdx.x * rows_per_cta, r_hi = min(rows, r_lo + rows_per_cta);
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int64_t r0 = r_lo; r0 < r_hi; r0 += 32) {
        __syncthreads();
        for (int idx = t; idx < 32 * 16; idx += kGramThreads) {
            const int rr = idx >> 4, c4 = idx & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + rr < r_hi) v = *reinterpret_cast<const float4 *>(P + (r0 + rr) * kQ + c4 * 4);
            *reinterpret_cast<float4 *>(&sP[rr][c4 * 4]) = v;
        }
        __syncthreads();
        if (active) {
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
                const float4 av = *reinterpret_cast<const float4 *>(&sP[rr][ty * 4]);
                const float4 bv = *reinterpret_cast<const float4 *>(&sP[rr][tx * 4]);
                const double am[4] = {av.x, av.y, av.z, av.w}, bm[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fma(am[i], bm[j], acc[i][j]);
            }
        }
    }
    if (!active) return;
    double *out = partial + (int64_t)blockIdx.x * kQ * kQ;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) out[(ty * 4 + i) * kQ + tx * 4 + j] = acc[i][j];
}

// Fixed summation order over the partials: eight contiguous ranges of them are summed by eight threads per entry and
// combined in range order.  Entries of a block below the diagonal are read from the mirrored block.
__global__ void __launch_bounds__(256)
panel_gram_reduce_kernel(const double *__restrict__ partial, int n_part, double *__restrict__ Gm) {
    __shared__ double seg_sum[8][32];
    const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;                       // 128 CTAs x 32 entries
    const int r = i / kQ, c = i % kQ;
    const int src = (r >> 2) <= (c >> 2) ? i : c * kQ + r;
    const int per = (n_part + 7) / 8, p_lo = seg * per, p_hi = min(n_part, p_lo + per);
    double s = 0.0;
#pragma unroll 4
    for (int p = p_lo; p < p_hi; ++p) s += partial[(int64_t)p * kQ * kQ + src];
    seg_sum[seg][lane] = s;
    __syncthreads();
    if (seg == 0) {
        double t = seg_sum[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += seg_sum[k][lane];
        Gm[i] = t;
    }
}

// Pout = P @ R ; optional column-wise |.|-max candidates (first index on ties) per CTA
template <bool ABSMAX>
__global__ void __launch_bounds__(256)
panel_matmul_kernel(const float *__restrict__ P, int64_t rows, const float *__restrict__ R, float *__restrict__ Pout,
                    float *__restrict__ cand_val, long long *__restrict__ cand_idx) {
    __shared__ __align__(16) float sR[kQ][kQ];
    __shared__ __align__(16) float sP[64][kQ + 4];
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    for (int idx = threadIdx.x; idx < kQ * kQ / 4; idx += 256)
        reinterpret_cast<float4 *>(&sR[0][0])[idx] = reinterpret_cast<const float4 *>(R)[idx];
    const int64_t r0 = (int64_t)blockIdx.x * 64;
    for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
        const int rr = idx >> 4, c4 = idx & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + rr < rows) v = *reinterpret_cast<const float4 *>(P + (r0 + rr) * kQ + c4 * 4);
        *reinterpret_cast<float4 *>(&sP[rr][c4 * 4]) = v;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int k = 0; k < kQ; ++k) {
        const float4 qv = *reinterpret_cast<const float4 *>(&sR[k][tx * 4]);
        const float qm[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float av = sP[ty * 4 + i][k];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av, qm[j], acc[i][j]);
        }
    }
    if (!ABSMAX) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t row = r0 + ty * 4 + i;
            if (row < rows)
                *reinterpret_cast<float4 *>(Pout + row * kQ + tx * 4) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    } else {
        __syncthreads();
        // reuse sP as the 64 x 64 product tile, then one thread per column scans rows in order
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) sP[ty * 4 + i][tx * 4 + j] = acc[i][j];
        __syncthreads();
        if (threadIdx.x < kQ) {
            float best = 0.f; long long bi = -1;
            const int nr = (int)min((int64_t)64, rows - r0);
            for (int rr = 0; rr < nr; ++rr) {
                const float v = sP[rr][threadIdx.x];
                if (bi < 0 || fabsf(v) > fabsf(best)) { best = v; bi = r0 + rr; }
            }
            cand_val[(int64_t)blockIdx.x * kQ + threadIdx.x] = best;
            cand_idx[(int64_t)blockIdx.x * kQ + threadIdx.x] = bi;
        }
    }
}

// one CTA per column; (|value| larger, else block index smaller) is associative, so the tree keeps the first-index rule
__global__ void __launch_bounds__(256)
colabsmax_reduce_kernel(const float *__restrict__ cand_val, int64_t n_cta, float *__restrict__ out_val) {
    __shared__ float s_val[256];
    __shared__ long long s_blk[256];
    const int j = blockIdx.x;
    float best = 0.f;
    long long best_b = -1;
    for (int64_t b = threadIdx.x; b < n_cta; b += 256) {       // ascending row blocks within a thread
        const float v = cand_val[b * kQ + j];
        if (best_b < 0 || fabsf(v) > fabsf(best)) { best = v; best_b = b; }
    }
    s_val[threadIdx.x] = best; s_blk[threadIdx.x] = best_b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const float v = s_val[threadIdx.x + o];
            const long long vb = s_blk[threadIdx.x + o];
            const float cur = s_val[threadIdx.x];
            const long long cb = s_blk[threadIdx.x];
            if (vb >= 0 && (cb < 0 || fabsf(v) > fabsf(cur) || (fabsf(v) == fabsf(cur) && vb < cb))) {
                s_val[threadIdx.x] = v; s_blk[threadIdx.x] = vb;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out_val[j] = s_val[0];
}

}  // namespace gc

using namespace gc;

extern "C" int gc_csr_to_counts(const float *data, const int32_t *indices, const int64_t *indptr, int64_t n_rows,
                                int64_t row0, int64_t G, uint16_t *counts, int64_t ld, int accumulate, int32_t *flags,
                                void *stream) {
    GC_REQUIRE(data && indices && indptr && counts && flags, "null pointer");
    GC_REQUIRE(G >= 1 && ld >= G && ld % 2 == 0 && ((uintptr_t)counts & 3) == 0, "bad counts layout");
    if (n_rows == 0) return 0;
    const unsigned grid = (unsigned)ceil_div<int64_t>(n_rows * 32, 256);
    if (accumulate)
        csr_to_counts_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(data, indices, indptr, n_rows, row0, G, counts, ld, flags);
    else
        csr_to_counts_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(data, indices, indptr, n_rows, row0, G, counts, ld, flags);
    return check_launch("csr_to_counts_kernel");
}

extern "C" int gc_dense_to_counts(const float *src, int64_t ld_src, int64_t n_rows, int64_t G, int64_t row0,
                                  uint16_t *counts, int64_t ld, int32_t *flags, void *stream) {
    GC_REQUIRE(src && counts && flags, "null pointer");
    if (n_rows == 0) return 0;
    GC_REQUIRE(n_rows <= 65535, "at most 65535 rows per call");
    dim3 grid((unsigned)ceil_div<int64_t>(G, 256), (unsigned)n_rows);
    dense_to_counts_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, ld_src, n_rows, G, row0, counts, ld, flags);
    return check_launch("dense_to_counts_kernel");
}

static int check_counts_layout(const uint16_t *counts, int64_t C, int64_t G, int64_t ld) {
    GC_REQUIRE(counts != nullptr, "null counts");
    GC_REQUIRE(C >= 1 && G >= 1, "empty matrix");
    GC_REQUIRE(ld >= G && ld % 8 == 0, "ld must be >= G and a multiple of 8");
    GC_REQUIRE(((uintptr_t)counts & 15) == 0, "counts must be 16-byte aligned");
    return 0;
}

namespace gc {
// Statistics the operator needs from the exact sums, in two launches instead of ~20 framework calls:
// rows: float32 copies, per-CTA (total, min, any zero);  finish: the reductions of those, float32 gene sums, any
// all-zero gene, and the clip proof  max_g x_max(g) / sqrt(r_min s_g / T)  >= 0.999 clip  (see ResidualOperator).
constexpr int kStatCtas = 148;
__global__ void __launch_bounds__(256)
count_stats_rows_kernel(const long long *__restrict__ row_sum, int64_t C, float *__restrict__ row_sum_f,
                        long long *__restrict__ part) {
    long long tot = 0, mn = 0x7fffffffffffffffLL, zero = 0;
    for (int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x; c < C; c += (int64_t)gridDim.x * 256) {
        const long long v = row_sum[c];
        row_sum_f[c] = (float)v;
        tot += v; mn = v < mn ? v : mn; zero |= (v == 0);
    }
    __shared__ long long s_tot[256], s_mn[256], s_zero[256];
    s_tot[threadIdx.x] = tot; s_mn[threadIdx.x] = mn; s_zero[threadIdx.x] = zero;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_tot[threadIdx.x] += s_tot[threadIdx.x + o];
            s_mn[threadIdx.x] = min(s_mn[threadIdx.x], s_mn[threadIdx.x + o]);
            s_zero[threadIdx.x] |= s_zero[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        part[blockIdx.x * 3 + 0] = s_tot[0]; part[blockIdx.x * 3 + 1] = s_mn[0]; part[blockIdx.x * 3 + 2] = s_zero[0];
    }
}

__global__ void __launch_bounds__(1024)
count_stats_finish_kernel(const long long *__restrict__ part, int n_part, const long long *__restrict__ col_sum,
                          const uint32_t *__restrict__ col_max, int64_t G, double clip, float *__restrict__ col_sum_f,
                          long long *__restrict__ head) {
    __shared__ long long s_tot, s_mn, s_zero;
    __shared__ double s_worst[1024];
    __shared__ int s_gzero[1024];
    if (threadIdx.x == 0) {
        long long tot = 0, mn = 0x7fffffffffffffffLL, zero = 0;
        for (int p = 0; p < n_part; ++p) {
            tot += part[p * 3]; mn = min(mn, part[p * 3 + 1]); zero |= part[p * 3 + 2];
        }
        s_tot = tot; s_mn = mn; s_zero = zero;
    }
    __syncthreads();
    const double total = (double)s_tot, r_min = (double)s_mn;
    double worst = 0.0;
    int gzero = 0;
    for (int64_t g = threadIdx.x; g < G; g += 1024) {
        const long long sg = col_sum[g];
        col_sum_f[g] = (float)sg;
        gzero |= (sg == 0);
        const double mu_min = r_min * (double)sg / total;
        double w = (double)col_max[g] / sqrt(mu_min);
        if (!(w == w)) w = INFINITY;                       // 0 / 0: an all-zero gene or cell (raised by the caller anyway)
        worst = fmax(worst, w);
    }
    s_worst[threadIdx.x] = worst; s_gzero[threadIdx.x] = gzero;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_worst[threadIdx.x] = fmax(s_worst[threadIdx.x], s_worst[threadIdx.x + o]);
            s_gzero[threadIdx.x] |= s_gzero[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        head[0] = s_tot; head[1] = s_gzero[0]; head[2] = s_zero; head[3] = s_worst[0] >= 0.999 * clip;
    }
}
}  // namespace gc

extern "C" int gc_count_sums(const uint16_t *counts, int64_t C, int64_t G, int64_t ld, unsigned long long *row_sum,
                             unsigned long long *col_sum, uint32_t *col_max, void *stream) {
    if (int rc = check_counts_layout(counts, C, G, ld)) return rc;
    GC_REQUIRE(row_sum && col_sum, "null pointer");
    // CTA width (8 columns per thread) with the least padding of the last column tile: a 2 560-column gene shard is five
    // 512-column tiles, not two 2 048-column ones
    int threads = 256;
    int64_t best_pad = -1;
    for (int t = 256; t >= 64; t -= 64) {
        const int64_t pad = ceil_div<int64_t>(ld, t * 8) * t * 8 - ld;
        if (best_pad < 0 || pad < best_pad) { best_pad = pad; threads = t; }
    }
    dim3 grid((unsigned)ceil_div<int64_t>(ld, threads * 8), (unsigned)ceil_div<int64_t>(C, kSumRows));
    count_sums_kernel<<<grid, threads, 0, as_stream(stream)>>>(counts, C, G, ld, row_sum, col_sum, col_max);
    return check_launch("count_sums_kernel");
}

extern "C" int gc_count_stats(const long long *row_sum, int64_t C, const long long *col_sum, const uint32_t *col_max,
                              int64_t G, double clip, float *row_sum_f, float *col_sum_f, long long *head,
                              long long *scratch, void *stream) {
    GC_REQUIRE(row_sum && col_sum && col_max && row_sum_f && col_sum_f && head && scratch, "null pointer");
    GC_REQUIRE(C >= 1 && G >= 1, "empty matrix");
    const int n_part = (int)std::min<int64_t>(kStatCtas, ceil_div<int64_t>(C, 256));
    count_stats_rows_kernel<<<n_part, 256, 0, as_stream(stream)>>>(row_sum, C, row_sum_f, scratch);
    if (int rc = check_launch("count_stats_rows_kernel")) return rc;
    count_stats_finish_kernel<<<1, 1024, 0, as_stream(stream)>>>(scratch, n_part, col_sum, col_max, G, clip, col_sum_f,
                                                                 head);
    return check_launch("count_stats_finish_kernel");
}

extern "C" int gc_pearson_residuals(const uint16_t *counts, int64_t C, int64_t G, int64_t ld, const float *row_sum_f,
                                    const float *col_sum_f, float total, float theta, float clip, float *Z,
                                    int64_t ldz, void *stream) {
    if (int rc = check_counts_layout(counts, C, G, ld)) return rc;
    GC_REQUIRE(row_sum_f && col_sum_f && Z && ldz >= G, "bad arguments");
    GC_REQUIRE(C <= 2147483647LL, "too many rows");
    const int64_t by = 65535;
    for (int64_t c0 = 0; c0 < C; c0 += by) {
        const int64_t nc = std::min<int64_t>(by, C - c0);
        dim3 grid((unsigned)ceil_div<int64_t>(G, 2048), (unsigned)nc);
        pearson_residuals_kernel<<<grid, 256, 0, as_stream(stream)>>>(counts + c0 * ld, nc, G, ld, row_sum_f + c0,
                                                                     col_sum_f, 1.0f / total, 1.0f / theta, clip,
                                                                     Z + c0 * ldz, ldz);
        if (int rc = check_launch("pearson_residuals_kernel")) return rc;
    }
    return 0;
}

// library-owned scratch of the column-sum pass, one buffer per (thread, device)
static thread_local double *g_col_partial_dev[kMaxDevices] = {};
static thread_local int64_t g_col_partial_cap_dev[kMaxDevices] = {};

extern "C" int gc_residual_sums(const uint16_t *counts, int64_t C, int64_t G, int64_t ld, const float *row_sum_f,
                                const float *col_sum_f, float total, float theta, float clip, double *zrow_sum,
                                double *zcol_sum, void *stream) {
    if (int rc = check_counts_layout(counts, C, G, ld)) return rc;
    GC_REQUIRE(row_sum_f && col_sum_f, "null pointer");
    const float inv_total = 1.0f / total, inv_theta = 1.0f / theta;
    if (zrow_sum) {
        residual_row_sums_kernel<<<(unsigned)ceil_div<int64_t>(C * 32, 256), 256, 0, as_stream(stream)>>>(
            counts, C, G, ld, row_sum_f, col_sum_f, inv_total, inv_theta, clip, zrow_sum);
        if (int rc = check_launch("residual_row_sums_kernel")) return rc;
    }
    if (zcol_sum) {
        const int64_t need = (int64_t)kColSlabs * G;
        double *&g_col_partial = g_col_partial_dev[current_device()];
        int64_t &g_col_partial_cap = g_col_partial_cap_dev[current_device()];
        if (need > g_col_partial_cap) {
            if (g_col_partial) GC_CUDA(cudaFree(g_col_partial));
            GC_CUDA(cudaMalloc(&g_col_partial, sizeof(double) * need));
            g_col_partial_cap = need;
        }
        dim3 grid((unsigned)ceil_div<int64_t>(G, 256), kColSlabs);
        residual_col_partials_kernel<<<grid, 256, 0, as_stream(stream)>>>(counts, C, G, ld, row_sum_f, col_sum_f,
                                                                         inv_total, inv_theta, clip, g_col_partial);
        if (int rc = check_launch("residual_col_partials_kernel")) return rc;
        residual_col_reduce_kernel<<<(unsigned)ceil_div<int64_t>(G, 256), 256, 0, as_stream(stream)>>>(g_col_partial, G,
                                                                                                      zcol_sum);
        if (int rc = check_launch("residual_col_reduce_kernel")) return rc;
    }
    return 0;
}

namespace gc {
int64_t simt_scratch_bytes(int64_t C, int64_t G) {
    // worst case over both directions
    const int s0 = choose_split(ceil_div<int64_t>(C, kTileM), G), s1 = choose_split(ceil_div<int64_t>(G, kTileM), C);
    const int64_t b0 = (int64_t)s0 * C * kQ * 4, b1 = (int64_t)s1 * G * kQ * 4;
    return (b0 > b1 ? b0 : b1) + 256;
}
}  // namespace gc

extern "C" int gc_rsvd_apply_simt(const gc_residual_matrix *M, int trans, const float *Qin, float *Y, int q,
                                  void *partials, void *stream) {
    GC_REQUIRE(M && Qin && Y && partials, "null pointer");
    if (int rc = check_counts_layout(M->counts, M->C, M->G, M->ld)) return rc;
    GC_REQUIRE(q >= 1 && q <= kQ, "q must be in [1, 64]");
    GC_REQUIRE(trans == 0 || trans == 1, "trans must be 0 or 1");
    ApplyArgs a;
    a.counts = M->counts; a.C = M->C; a.G = M->G; a.ld = M->ld;
    a.row_sum_f = M->row_sum_f; a.col_sum_f = M->col_sum_f;
    a.inv_total = 1.0f / M->total; a.inv_theta = 1.0f / M->theta; a.clip = M->clip;
    a.row_shift = M->row_shift; a.col_shift = M->col_shift;
    a.Qin = Qin; a.partials = (float *)partials;
    a.n_out = trans == 0 ? M->C : M->G;
    a.n_k = trans == 0 ? M->G : M->C;
    const int64_t n_tiles = ceil_div<int64_t>(a.n_out, kTileM);
    a.n_split = choose_split(n_tiles, a.n_k);
    a.k_per_split = ceil_div<int64_t>(ceil_div<int64_t>(a.n_k, a.n_split), kTileK) * kTileK;
    dim3 grid((unsigned)n_tiles, (unsigned)a.n_split);
    prof_begin(kProfRsvdPass, as_stream(stream));
    if (trans == 0) rsvd_apply_simt_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(a);
    else rsvd_apply_simt_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(a);
    prof_end(kProfRsvdPass, as_stream(stream));
    if (int rc = check_launch("rsvd_apply_simt_kernel")) return rc;
    const int64_t n4 = a.n_out * kQ / 4;
    rsvd_reduce_partials_kernel<<<(unsigned)ceil_div<int64_t>(n4, 256), 256, 0, as_stream(stream)>>>(
        a.partials, n4, a.n_split, Y);
    return check_launch("rsvd_reduce_partials_kernel");
}

extern "C" int64_t gc_gram_scratch_bytes(int64_t rows) {
    // upper bound of the partial count for every row count <= rows (a shorter panel may use more, smaller slices)
    const int64_t n_part = std::min<int64_t>(kGramMaxCtas, ceil_div<int64_t>(rows, 32));
    const int64_t n_cta64 = ceil_div<int64_t>(rows, 64);
    const int64_t a = n_part * kQ * kQ * 8, b = n_cta64 * kQ * (4 + 8);
    return (a > b ? a : b) + 256;
}

extern "C" int gc_panel_gram(const float *P, int64_t rows, double *Gm, void *scratch, void *stream) {
    GC_REQUIRE(P && Gm && scratch && rows >= 1, "bad arguments");
    const int64_t per = gram_rows_per_cta(rows);
    const int n_part = (int)ceil_div<int64_t>(rows, per);
    panel_gram_partial_kernel<<<n_part, kGramThreads, 0, as_stream(stream)>>>(P, rows, per, (double *)scratch);
    if (int rc = check_launch("panel_gram_partial_kernel")) return rc;
    panel_gram_reduce_kernel<<<kQ * kQ / 32, 256, 0, as_stream(stream)>>>((const double *)scratch, n_part, Gm);
    return check_launch("panel_gram_reduce_kernel");
}

extern "C" int gc_panel_matmul(const float *P, int64_t rows, const float *R, float *Pout, void *stream) {
    GC_REQUIRE(P && R && Pout && rows >= 1, "bad arguments");
    panel_matmul_kernel<false><<<(unsigned)ceil_div<int64_t>(rows, 64), 256, 0, as_stream(stream)>>>(P, rows, R, Pout,
                                                                                                     nullptr, nullptr);
    return check_launch("panel_matmul_kernel");
}

extern "C" int gc_panel_colabsmax(const float *P, int64_t rows, const float *R, float *out_val, void *scratch,
                                  void *stream) {
    GC_REQUIRE(P && R && out_val && scratch && rows >= 1, "bad arguments");
    const int64_t n_cta = ceil_div<int64_t>(rows, 64);
    float *cand_val = (float *)scratch;
    long long *cand_idx = (long long *)((char *)scratch + ((n_cta * kQ * 4 + 255) / 256) * 256);
    panel_matmul_kernel<true><<<(unsigned)n_cta, 256, 0, as_stream(stream)>>>(P, rows, R, nullptr, cand_val, cand_idx);
    if (int rc = check_launch("panel_matmul_kernel<absmax>")) return rc;
    colabsmax_reduce_kernel<<<kQ, 256, 0, as_stream(stream)>>>(cand_val, n_cta, out_val);
    return check_launch("colabsmax_reduce_kernel");
}
