// High-confidence-cell stage of GeneClust-ps (scGeneClust/tl/confidence.py:30-148): the cells x cells co-membership
// "frequency matrix".  The reference fits a Gaussian mixture on each of R prefixes of the cell PCs and, per fit, builds a
// dense C x C matrix in a Python loop over cells (confidence.py:111-114):
//     co[i, j] = 1  for j > i  iff  proba_i >= p  and  proba_j > p  and  label_i == label_j
// then sums the R matrices (confidence.py:60).  Here: one kernel, one byte per cell pair, the R fits folded in registers.
// Integer / comparison work on the mixture's labels and maximum posteriors: bit-exact against the reference
// (tests/golden/confidence_small.npz is the output of the reference's own _compute_cell_co_membership).
#include "common.cuh"

namespace gc {

constexpr int kCoTile = 64;          // cells per tile edge
constexpr int kCoMaxRuns = 32;

// key[r][c] = label + 1 if the cell passes the row test (proba >= p), 0 otherwise; colkey likewise for the column test
// (proba > p).  co-membership of (i, j) in run r  <=>  rowkey[r][i] != 0 && rowkey[r][i] == colkey[r][j].
__global__ void comembership_keys_kernel(const int32_t *__restrict__ labels, const double *__restrict__ probas, int64_t n,
                                         double p, int32_t *__restrict__ rowkey, int32_t *__restrict__ colkey) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double pr = probas[i];
    const int32_t k = labels[i] + 1;
    rowkey[i] = pr >= p ? k : 0;
    colkey[i] = pr > p ? k : 0;
}

__global__ void __launch_bounds__(256)
comembership_counts_kernel(const int32_t *__restrict__ rowkey, const int32_t *__restrict__ colkey, int64_t C, int R,
                           uint8_t *__restrict__ freq) {
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj < ti) return;                                   // tiles strictly below the diagonal hold zeros (pre-zeroed)
    __shared__ int32_t srow[kCoMaxRuns][kCoTile];
    __shared__ int32_t scol[kCoMaxRuns][kCoTile];
    const int64_t i0 = (int64_t)ti * kCoTile, j0 = (int64_t)tj * kCoTile;
    for (int e = threadIdx.x; e < R * kCoTile; e += blockDim.x) {
        const int r = e / kCoTile, c = e % kCoTile;
        srow[r][c] = i0 + c < C ? rowkey[(int64_t)r * C + i0 + c] : 0;
        scol[r][c] = j0 + c < C ? colkey[(int64_t)r * C + j0 + c] : -1;
    }
    __syncthreads();
    // thread = (row group of 4 rows, column): 16 x 16 threads, each 4 rows x 4 columns (columns strided by 16: coalesced)
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int li = ty * 4 + a;
        const int64_t i = i0 + li;
        if (i >= C) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int lj = tx + 16 * b;
            const int64_t j = j0 + lj;
            if (j >= C || j <= i) continue;
            int cnt = 0;
            for (int r = 0; r < R; ++r) {
                const int32_t k = srow[r][li];
                cnt += (k != 0 && k == scol[r][lj]);
            }
            freq[i * C + j] = (uint8_t)cnt;
        }
    }
}

}  // namespace gc

using namespace gc;

extern "C" int gc_comembership_counts(const int32_t *labels, const double *probas, int64_t C, int R, double p,
                                      uint8_t *freq, int32_t *scratch, void *stream) {
    GC_REQUIRE(labels && probas && freq && scratch, "null pointer");
    GC_REQUIRE(C >= 1 && R >= 1 && R <= kCoMaxRuns, "1 <= runs <= 32");
    cudaStream_t st = as_stream(stream);
    int32_t *rowkey = scratch, *colkey = scratch + (int64_t)R * C;
    const int64_t n = (int64_t)R * C;
    comembership_keys_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, st>>>(labels, probas, n, p, rowkey, colkey);
    if (int rc = check_launch("comembership_keys_kernel")) return rc;
    GC_CUDA(cudaMemsetAsync(freq, 0, (size_t)C * C, st));
    const unsigned tiles = (unsigned)ceil_div<int64_t>(C, kCoTile);
    GC_REQUIRE(tiles <= 65535, "too many cells for the dense frequency matrix");
    comembership_counts_kernel<<<dim3(tiles, tiles), 256, 0, st>>>(rowkey, colkey, C, R, freq);
    return check_launch("comembership_counts_kernel");
}
