/* C twin of oracle/synth.py::synth_counts (TEST/BENCH INFRASTRUCTURE, see oracle/__init__.py).
 *
 * The numpy twin needs ~0.1 s per cell at 20 000 genes (its masked while-loop runs once per count level), so it
 * cannot generate the full-size inputs of the CPU reference arm (50 000 x 20 000) or of the 20 000-cell subsample
 * harness.  This file restates the same generator in plain C with OpenMP over cells:
 * integer splitmix hashing + IEEE float64 mul/div/add in the same order, compiled with -ffp-contract=off, so the
 * counts are bit-identical to oracle/synth.py (checked in tests/test_host_logic.py) and to the device generator
 * gc_synth_counts (checked in tests/test_gpu_fullsize.py).  Nothing of the reference is restated here: its loaders
 * download real data sets (scGeneClust/_utils.py:17-74); SURVEY.md section 8(d) defines this synthetic stream.
 *
 * Built by oracle/build_c.py into oracle/_build/liboracle_synth.so; loaded with ctypes by oracle/synth.py.
 */
#include <stdint.h>

static inline uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}

static inline uint64_t hash3(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (stream + 1);
    z = mix64(z ^ (a * 0xBF58476D1CE4E5B9ULL));
    z = mix64(z ^ (b * 0x94D049BB133111EBULL));
    return z;
}

/* counts[(c - c0) * ld + (g - g0)] for cells [c0, c0+nc) x genes [g0, g0+ng) */
void oracle_synth_counts(uint64_t seed, int64_t c0, int64_t nc, int64_t g0, int64_t ng, const double *gene_mean,
                         const double *type_fc, int n_types, const int32_t *gene_module, const double *size_levels,
                         int n_size, const double *module_levels, int n_modlev, uint16_t *counts, int64_t ld) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t cl = 0; cl < nc; ++cl) {
        const uint64_t c = (uint64_t)(c0 + cl);
        const int t = (int)(hash3(seed, 1, c, 0) % (uint64_t)n_types);
        const int sc = (int)(hash3(seed, 2, c, 0) % (uint64_t)n_size);
        uint16_t *row = counts + cl * ld;
        for (int64_t gl = 0; gl < ng; ++gl) {
            const uint64_t g = (uint64_t)(g0 + gl);
            double mean = gene_mean[g] * type_fc[g * n_types + t];
            mean = mean * size_levels[sc];
            const int mod = gene_module[g];
            if (mod >= 0) mean = mean * module_levels[hash3(seed, 3, c, (uint64_t)mod) % (uint64_t)n_modlev];
            const double u = (double)(hash3(seed, 4, c, g) >> 11) * 1.1102230246251565e-16; /* 2^-53 */
            /* negative binomial, r = 2: pmf(k) = (k+1) p^2 q^k, inversion by sequential search */
            const double denom = 2.0 + mean;
            const double p = 2.0 / denom, q = mean / denom;
            double pmf = p * p, cdf = pmf;
            int k = 0;
            while (u > cdf && k < 65535 && pmf > 0.0) {
                ++k;
                pmf = (pmf * q) * ((double)(k + 1) / (double)k);
                cdf = cdf + pmf;
            }
            row[gl] = (uint16_t)k;
        }
    }
}

int oracle_synth_abi(void) { return 1; }
