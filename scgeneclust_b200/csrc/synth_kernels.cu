// Counter-based synthetic negative-binomial count generator (bench/test harness).
// Bit-identical twin of oracle/synth.py: integer hashing + IEEE float64 mul/div/add only (-fmad=false).
#include "common.cuh"

namespace gc {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
__host__ __device__ __forceinline__ uint64_t hash3(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (stream + 1);
    z = mix64(z ^ (a * 0xBF58476D1CE4E5B9ULL));
    z = mix64(z ^ (b * 0x94D049BB133111EBULL));
    return z;
}

__global__ void __launch_bounds__(256)
synth_counts_kernel(uint64_t seed, int64_t c0, int64_t nc, int64_t g0, int64_t ng,
                    const double *__restrict__ gene_mean, const double *__restrict__ type_fc, int n_types,
                    const int32_t *__restrict__ gene_module, const double *__restrict__ size_levels, int n_size,
                    const double *__restrict__ module_levels, int n_modlev, uint16_t *__restrict__ counts, int64_t ld) {
    const int64_t gl = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // local gene
    const int64_t cl = blockIdx.y;                                       // local cell
    if (gl >= ng || cl >= nc) return;
    const uint64_t c = (uint64_t)(c0 + cl), g = (uint64_t)(g0 + gl);
    const int t = (int)(hash3(seed, 1, c, 0) % (uint64_t)n_types);
    const int sc = (int)(hash3(seed, 2, c, 0) % (uint64_t)n_size);
    double mean = gene_mean[g] * type_fc[g * n_types + t];
    mean = mean * size_levels[sc];
    const int mod = gene_module[g];
    if (mod >= 0) mean = mean * module_levels[hash3(seed, 3, c, (uint64_t)mod) % (uint64_t)n_modlev];
    const double u = (double)(hash3(seed, 4, c, g) >> 11) * 1.1102230246251565e-16;  // 2^-53
    // negative binomial with r = 2 (phi = 0.5): pmf(k) = (k+1) p^2 q^k, inversion by sequential search
    const double denom = 2.0 + mean;
    const double p = 2.0 / denom, q = mean / denom;
    double pmf = p * p, cdf = pmf;
    int k = 0;
    while (u > cdf && k < 65535 && pmf > 0.0) {
        ++k;
        pmf = (pmf * q) * ((double)(k + 1) / (double)k);
        cdf = cdf + pmf;
    }
    counts[cl * ld + gl] = (uint16_t)k;
}

}  // namespace gc

using namespace gc;

extern "C" int gc_synth_counts(uint64_t seed, int64_t c0, int64_t nc, int64_t g0, int64_t ng, const double *gene_mean,
                               const double *type_fc, int n_types, const int32_t *gene_module,
                               const double *size_levels, int n_size, const double *module_levels, int n_modlev,
                               int n_modules, uint16_t *counts, int64_t ld, void *stream) {
    GC_REQUIRE(gene_mean && type_fc && gene_module && size_levels && module_levels && counts, "null pointer");
    GC_REQUIRE(nc >= 1 && ng >= 1 && ld >= ng && n_types >= 1 && n_size >= 1 && n_modlev >= 1, "bad shape");
    (void)n_modules;
    for (int64_t r0 = 0; r0 < nc; r0 += 65535) {
        const int64_t nr = std::min<int64_t>(65535, nc - r0);
        dim3 grid((unsigned)ceil_div<int64_t>(ng, 256), (unsigned)nr);
        synth_counts_kernel<<<grid, 256, 0, as_stream(stream)>>>(seed, c0 + r0, nr, g0, ng, gene_mean, type_fc, n_types,
                                                                gene_module, size_levels, n_size, module_levels,
                                                                n_modlev, counts + r0 * ld, ld);
        if (int rc = check_launch("synth_counts_kernel")) return rc;
    }
    return 0;
}
