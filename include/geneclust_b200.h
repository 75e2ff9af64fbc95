/* geneclust_b200.h - C ABI of libgeneclust_b200.so (sm_100a).
 *
 * Drop-in boundary for the gene-grouping hot path of GeneClust (ToryDeng/scGeneClust).  The reference has no
 * FFI layer: its boundary is the Python stage functions that mutate an AnnData.  Each
 * entry point below replaces the third-party native arithmetic one of those stage functions runs today; the
 * reference line it replaces is cited on every declaration (paths relative to the reference checkout).  The
 * Python binding a maintainer would add is shown in INTEGRATION.md; this repo's own host side
 * (scgeneclust_b200/_lib.py) binds exactly these symbols with ctypes.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer (caller-owned, e.g. torch tensor storage) unless its name ends in
 *     `_host`.  Outputs are preallocated by the caller.  No ownership is transferred.
 *   - `stream` is a cudaStream_t passed as void*.  All work is enqueued on it; calls do not synchronise unless
 *     stated.  Calls are re-entrant per (device, stream).
 *   - Return value: 0 on success, non-zero on error; gc_last_error() returns a thread-local message.
 *   - Matrices are row-major.  "counts" is the cells x genes raw-count matrix stored as uint16.
 */
#ifndef GENECLUST_B200_H
#define GENECLUST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GC_ABI_VERSION 1

int gc_abi_version(void);
const char *gc_last_error(void);
/* Number of kernels this library has launched in the calling process (all streams); bench.py's gpu_launches. */
long long gc_launch_count(void);

/* Device-time measurement with CUDA events on the launching stream (bench.py's roofline block).  While enabled, the
 * entry points that launch a dominant kernel bracket exactly that kernel with an event pair:
 *   slot 0 = the randomized-SVD pass kernel (gc_rsvd_apply, gc_rsvd_apply_simt), slot 1 = the KSG pair kernel
 *   (gc_mi_cc_pairs).  gc_profile_collect synchronises the recorded events, returns their summed duration and
 * count (host pointers) and clears the slot. */
int gc_profile_enable(int on);
int gc_profile_collect(int slot, double *total_ms_host, long long *launches_host);

/* Measured fp64 issue rate of this GPU (DFMA lane-operations per second, 8 independent chains per thread, all SMs):
 * the roofline denominator bench.py uses for the kNN mutual-information kernels, which are fp64 / integer issue bound
 * (SURVEY.md section 8d).  scratch_dev: 8 bytes of device memory.  Synchronises the stream. */
int gc_fp64_probe(double *ops_per_s_host, double *scratch_dev, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Ingest.  Replaces the host densification at scGeneClust/_model.py:123-124 and the integer check at
 * scGeneClust/_validation.py:68-70.
 * ------------------------------------------------------------------------------------------------------- */
/* CSR (float32 data, int32 indices, int64 indptr, `n_rows` rows starting at global row `row0` of the dense
 * matrix, G columns) -> dense uint16 counts[C x ld].  flags (4 x int32, zeroed by the caller) are set to 1 when:
 * [0] a value is not a non-negative integer, [1] a value or a sum of duplicates exceeds 65535, [2] a column index is
 * outside [0, G), [3] (accumulate == 0 only) a row is not canonical, i.e. its indices are not strictly increasing.
 * accumulate == 0: plain stores, correct for canonical rows; accumulate != 0: duplicate entries are summed, as scipy's
 * toarray() does for the reference (scGeneClust/_model.py:124).  The destination rows must be zeroed by the caller. */
int gc_csr_to_counts(const float *data, const int32_t *indices, const int64_t *indptr, int64_t n_rows,
                     int64_t row0, int64_t G, uint16_t *counts, int64_t ld, int accumulate, int32_t *flags,
                     void *stream);
/* Dense float32 block (n_rows x G, leading dimension ld_src) -> uint16 counts rows [row0, row0+n_rows). */
int gc_dense_to_counts(const float *src, int64_t ld_src, int64_t n_rows, int64_t G, int64_t row0,
                       uint16_t *counts, int64_t ld, int32_t *flags, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * a1  Analytic Pearson residuals.  Replaces scanpy normalize_pearson_residuals called at
 * scGeneClust/pp/_preprocessing.py:29.  The residual matrix is never required to be materialised: the
 * statistics below fully determine Z[c,g] = clip((x - mu)/sqrt(mu + mu^2/theta), +-sqrt(C)), mu = r_c*s_g/T.
 * ------------------------------------------------------------------------------------------------------- */
/* Exact integer row sums (per cell, C), column sums (per gene, G) of counts[C x G]; outputs are uint64 and
 * are ACCUMULATED into (zero them first), so a gene-sharded caller can all-reduce row sums. */
/* col_max (G x uint32, zeroed by the caller, or NULL): the largest count of every gene. */
int gc_count_sums(const uint16_t *counts, int64_t C, int64_t G, int64_t ld, unsigned long long *row_sum,
                  unsigned long long *col_sum, uint32_t *col_max, void *stream);
/* Derived statistics of the exact sums (row_sum already all-reduced on gene-sharded input): float32 copies, and
 * head[4] (int64) = { total count, any all-zero gene, any all-zero cell, clip proof failed: some residual MAY reach
 * +clip }.  scratch: 3 * 148 int64.  Replaces the reference's X.sum(0) / X.sum(1) / X.sum() bookkeeping inside
 * scanpy's normalize_pearson_residuals (reached from scGeneClust/pp/_preprocessing.py:29). */
int gc_count_stats(const long long *row_sum, int64_t C, const long long *col_sum, const uint32_t *col_max, int64_t G,
                   double clip, float *row_sum_f, float *col_sum_f, long long *head, long long *scratch, void *stream);
/* Materialise Z (float32, C x G, leading dimension ldz) - used by the ps path, which needs adata.X = residuals
 * (scGeneClust/tl/information.py:46,119), and by tests. */
int gc_pearson_residuals(const uint16_t *counts, int64_t C, int64_t G, int64_t ld, const float *row_sum_f,
                         const float *col_sum_f, float total, float theta, float clip, float *Z, int64_t ldz,
                         void *stream);
/* Per-cell and per-gene SUMS of Z in float64 (accumulated into the outputs; zero them first).  The PCA
 * centring vectors of sklearn PCA (decomposition/_pca.py:733-735) are these sums divided by G resp. C. */
int gc_residual_sums(const uint16_t *counts, int64_t C, int64_t G, int64_t ld, const float *row_sum_f,
                     const float *col_sum_f, float total, float theta, float clip, double *zrow_sum,
                     double *zcol_sum, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * a2  Gene-by-cell embedding.  Replaces the BLAS passes of sklearn randomized SVD
 * (utils/extmath.py:377-383, :606) reached from sc.pp.pca at scGeneClust/pp/_preprocessing.py:52,54.
 * M = Z - row_shift[c] - col_shift[g] is applied implicitly (fused residual producer, implicit centring).
 *   trans = 0:  Y[C x q] = M  @ Qin[G x q]      (contract over genes)
 *   trans = 1:  Y[G x q] = M^T @ Qin[C x q]     (contract over cells)
 * q <= 64; Qin/Y have leading dimension 64 (columns q..63 are ignored / written as zero).
 * `partials` is scratch of gc_rsvd_scratch_bytes(); results are deterministic (fixed reduction order).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct {
    const uint16_t *counts; int64_t C, G, ld;
    const float *row_sum_f, *col_sum_f; float total, theta, clip;
    const float *row_shift;   /* C or NULL */
    const float *col_shift;   /* G or NULL */
    int32_t upper_clip_inactive;   /* != 0: the caller has PROVED that no residual reaches +clip (gc_count_sums col_max:
                                    * max_g x_max(g) / sqrt(r_min s_g / T) < clip), so gc_rsvd_apply may skip the test */
} gc_residual_matrix;

int64_t gc_rsvd_scratch_bytes(int64_t C, int64_t G);
int gc_rsvd_apply(const gc_residual_matrix *M, int trans, const float *Qin, float *Y, int q, void *partials,
                  void *stream);
/* The previous tensor-core implementation of the same pass (both operands from shared memory, "SS mode"; gc_rsvd_apply
 * keeps the residual operand in tensor memory and feeds the count tiles by TMA): kept for A/B timing and as a second
 * tcgen05 implementation the tests compare with. */
int gc_rsvd_apply_ss(const gc_residual_matrix *M, int trans, const float *Qin, float *Y, int q, void *partials,
                     void *stream);
/* The SIMT (FFMA, exact float32 products) implementation of the same pass; kept as the on-device cross-check of
 * the tensor-core kernels. */
int gc_rsvd_apply_simt(const gc_residual_matrix *M, int trans, const float *Qin, float *Y, int q, void *partials,
                       void *stream);
/* The test matrix Omega of the randomized solver, drawn on the device exactly as scikit-learn draws it on the host
 * (utils/extmath.py:323-334: random_state.normal(size=(rows, q)) cast to float32): legacy numpy RandomState stream =
 * MT19937 + polar-method gauss.  mt_state: the 624 MT19937 key words of RandomState(seed).get_state()[1] (device),
 * mt_pos = get_state()[2].  Writes panel[r*ld + c] for c < q (other columns untouched).  status[0] |= 2 if the
 * pre-generated word stream held too few accepted polar attempts (caller should fall back / retry; probability ~0).
 * scratch: gc_mt19937_normal_scratch_bytes(rows, q), 256-byte aligned. */
int64_t gc_mt19937_normal_scratch_bytes(int64_t rows, int q);
int gc_mt19937_normal_panel(const uint32_t *mt_state, int mt_pos, int64_t rows, int q, float *panel, int ld,
                            void *scratch, int32_t *status, void *stream);
/* Small panel algebra on tall-skinny panels (rows x 64, float32):
 *   gram:  Gm[64 x 64] (float64) = P^T P   (deterministic two-stage reduction; scratch >= gc_gram_scratch_bytes)
 *   right-multiply: Pout = P @ R[64 x 64] (R float32, row-major) */
int64_t gc_gram_scratch_bytes(int64_t rows);
int gc_panel_gram(const float *P, int64_t rows, double *Gm, void *scratch, void *stream);
int gc_panel_matmul(const float *P, int64_t rows, const float *R, float *Pout, void *stream);
/* Column-wise index of the max-|.| entry of P[rows x 64] @ R (for svd_flip, utils/extmath.py:975-981):
 * out_val[j] = the signed entry of largest magnitude in column j (first index on ties). */
int gc_panel_colabsmax(const float *P, int64_t rows, const float *R, float *out_val, void *scratch, void *stream);
/* q x q float64 algebra on one CTA (replaces LAPACK getrf/geqrf/gesdd of utils/extmath.py:377-386, :606-619):
 *   gc_chol_inv: Gm = L L^T (leading q x q of a 64 x 64 row-major matrix); Rinv = (L^-1)^T as float32 64 x 64, so
 *                P @ Rinv is orthonormal when Gm = P^T P.  status[0] |= 1 on a non-positive pivot.
 *   gc_sym_eig:  eigenvalues (descending, evals[64]) and eigenvectors (columns of U32 / optional U64) by cyclic Jacobi.
 *   gc_scale_cols: R[:, j] *= sign(sign_src[j]) * scale[j] (either may be NULL) for j < q, zero for j >= q. */
int gc_chol_inv(const double *Gm, int q, float *Rinv, int32_t *status, void *stream);
int gc_sym_eig(const double *Gm, int q, double *evals, float *U32, double *U64, void *stream);
int gc_scale_cols(float *R, const float *sign_src, const double *scale, int q, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * a3/a4  GeneClust-fast gene k-means.  Replaces sklearn's Cython/OpenMP/sgemm kernels reached from
 * MiniBatchKMeans.fit_predict at scGeneClust/tl/cluster.py:66-70 and paired_distances at :101.
 * X is n x d float32 (d <= 64), centers K x d float32.
 * ------------------------------------------------------------------------------------------------------- */
/* labels[s] = argmin_j (|c_j|^2 - 2 x.c_j) with first-index ties (cluster/_k_means_lloyd.pyx:201-211);
 * sqdist[s] = |x - c_label|^2 summed in the 4-way order of _euclidean_dense_dense (_k_means_common.pyx:16-43).
 * If `gather` != NULL sample s is row gather[s] of X. */
int gc_kmeans_assign(const float *X, int64_t ldx, const int32_t *gather, int64_t n, const float *centers,
                     int K, int d, int32_t *labels, float *sqdist, void *stream);
/* One dense minibatch centre update with unit sample weights (cluster/_k_means_minibatch.pyx:60-108):
 * sequential float32 accumulation in sample order, weight_sums updated in place. */
int gc_kmeans_minibatch_update(const float *X, int64_t ldx, const int32_t *gather, int64_t n,
                               const int32_t *labels, const float *centers_old, float *centers_new,
                               float *weight_sums, int K, int d, void *stream);
/* Squared euclidean distances of every row of X to `ncand` candidate rows, computed in float64 and rounded
 * to float32, clamped at 0 (metrics/pairwise.py _euclidean_distances float32 upcast path) - the k-means++
 * building block (cluster/_kmeans.py:231-262).  out is ncand x n. */
int gc_sqdist_to_rows(const float *X, int64_t ldx, int64_t n, int d, const int32_t *cand, int ncand,
                      float *out, void *stream);
/* k-means++ seeding (cluster/_kmeans.py:180-275), device resident: the K-1 greedy rounds are enqueued back to back.
 * The host supplies the first centre id (its RandomState.choice draw) and the (K-1) x n_trials uniforms it drew
 * (their number does not depend on the data).  Potentials are float64 sums rounded to float32; the cumulative sum
 * searched by np.searchsorted is the sequential float32 np.cumsum.  ws: gc_kmeanspp_scratch_bytes(n, n_trials). */
int64_t gc_kmeanspp_scratch_bytes(int64_t n, int n_trials);
int gc_kmeanspp_init(const float *X, int64_t ldx, int64_t n, int d, int K, int32_t first_center,
                     const double *uniforms, int n_trials, void *ws, float *centers_out, int32_t *indices_out,
                     void *stream);
/* idx[i] = np.searchsorted(cdf, u[i], side='right') - the index stream of RandomState.choice(n, m, p=p) given its
 * uniform draws u (MiniBatchKMeans mini-batch sampling, cluster/_kmeans.py:2169-2174).  cdf: float64[n] ascending. */
int gc_searchsorted_right(const double *cdf, int64_t n, const double *u, int m, int32_t *idx, void *stream);
/* out[0] = float64 sum of a float32 vector in a fixed order (batch inertia, cluster/_k_means_common.pyx:94-124). */
int gc_sum_f32(const float *v, int64_t n, double *out, void *stream);
/* One step of MiniBatchKMeans.fit (sklearn cluster/_kmeans.py:2169-2196, reached from scGeneClust/tl/cluster.py:70) as one
 * call: idx = searchsorted(cdf, u, 'right') (the host's uniform draws -> batch rows), assignment of X[idx] to `centers`,
 * *inertia = sum of the squared distances (float64), running-mean update into centers_new / counts. */
int gc_kmeans_minibatch_step(const float *X, int64_t ldx, int64_t n, int d, int K, const double *cdf, const double *u,
                             int batch, int32_t *idx, int32_t *labels_b, float *sq_b, const float *centers,
                             float *centers_new, float *counts, double *inertia, void *stream);
/* Random reassignment of starved centres (cluster/_kmeans.py:1655-1682): centers_new[clusters[r]] = X[idx[picks[r]]],
 * counts[clusters[r]] = min_count, r < n_reassign (clusters / picks: device int32). */
int gc_kmeans_reassign(const float *X, int64_t ldx, int d, const int32_t *idx, const int32_t *clusters,
                       const int32_t *picks, int n_reassign, float min_count, float *centers_new, float *counts,
                       void *stream);
/* dist[s] = |x_s - centers[labels[s]]|_2 (paired_distances, scGeneClust/tl/cluster.py:101). */
int gc_paired_dist(const float *X, int64_t ldx, int64_t n, int d, const float *centers, const int32_t *labels,
                   float *dist, void *stream);
/* closeness[s] = 1 - minmax_scale(dist within the cluster of s) in float32, sklearn's `x * scale + min` form with
 * scale = 1 / range (range < 10 eps -> 1) (scGeneClust/tl/cluster.py:102-104).  scratch: 2*K int32. */
int gc_closeness(const float *dist, const int32_t *labels, int64_t n, int K, int32_t *scratch, float *closeness,
                 void *stream);
/* Per-cluster size and the FIRST index of the maximum / minimum of a non-negative float32 value (the closeness): what the
 * pandas groupby(...).nlargest(1) / .nsmallest(1) of scGeneClust/tl/selection.py:57-59 pick.  labels in [0, K).
 * out: 3 K int32 = counts | first arg-max | first arg-min (0x7fffffff for an empty cluster); scratch: 2 K int32. */
int gc_cluster_extremes(const float *value, const int32_t *labels, int64_t n, int K, int32_t *out, int32_t *scratch,
                        void *stream);

/* a5  Binomial deviance of the singleton-cluster genes (post-hoc filtering): compute_deviance of
 * scGeneClust/tl/selection.py:82-89 on the float32 residual columns X (C x m, row-major, contiguous), with numpy's
 * float32 operation and summation order (pairwise / row-after-row); out[m] float32.
 * scratch: gc_binomial_deviance_scratch_bytes(C, m), 256-byte aligned. */
int64_t gc_binomial_deviance_scratch_bytes(int64_t C, int m);
int gc_binomial_deviance(const float *X, int64_t C, int m, float *out, void *scratch, void *stream);
/* The building block of the above, exposed for tests: out[j] = numpy's a.sum(0)[j] for a C-ordered float32 (C, m > 1)
 * array, i.e. the sequential float32 chain a[0,j] + a[1,j] + ... (np.nansum at scGeneClust/tl/selection.py:86-87 after
 * the NaN replacement), evaluated piece-parallel and bit-exact for m <= 64.  replays (m int32, or NULL): how many of the
 * ceil(C / 256) pieces had to be replayed add by add.  scratch: gc_binomial_deviance_scratch_bytes(C, m). */
int gc_sum_axis0_f32(const float *a, int64_t C, int m, float *out, int32_t *replays, void *scratch, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * a6/a7/a8  kNN mutual information (GeneClust-ps).  Replaces sklearn mutual_info_classif /
 * mutual_info_regression (feature_selection/_mutual_info.py:20-153, :294-315) called per gene / per pair /
 * per MST edge at scGeneClust/tl/information.py:24, :63, :96.  k = 3 neighbours.
 * ------------------------------------------------------------------------------------------------------- */
/* Input preparation of `_estimate_mi` (scale + 1e-10 noise), bit-exact incl. numpy's pairwise summation.
 * Z: float32, row-major, `ldz`; column g of Z restricted to rows order[seg_off[c] .. seg_off[c+1]) is one
 * slice.  For every (gene g < G, segment c < n_seg):
 *   XR[g*n_tot + r] = feature-role (float64 path) value, YR[...] = target-role (float32 path) value,
 * r running over seg_off[c]..seg_off[c+1].  xi1/xi2: noise vectors laid out like the segments (n_tot).
 * YR/xi2 may be NULL. */
int gc_mi_prepare(const float *Z, int64_t ldz, const int32_t *order, const int32_t *seg_off, int n_seg,
                  int64_t G, int64_t n_tot, const double *xi1, const double *xi2, double *XR, double *YR,
                  void *stream);
/* KSG MI for gene pairs in itertools.combinations order, condensed index p in [p_begin, p_end):
 * pair (i<j) uses XR[i, :n] and YR[j, :n] (n <= 64).  psi[m] = digamma(m), m = 0..n (psi[0] unused).
 * mi_condensed[p - p_begin] always written; if dense != NULL also dense[i*N+j] = dense[j*N+i] = mi.
 * nx_out/ny_out (uint8, (p_end-p_begin) x n) optional: the integer neighbour counts. */
int gc_mi_cc_pairs(const double *XR, const double *YR, int64_t N, int n, int64_t ldr, const double *psi,
                   int64_t p_begin, int64_t p_end, double *mi_condensed, double *dense, uint8_t *nx_out,
                   uint8_t *ny_out, void *stream);
/* Same results from sorted 1-D orders prepared once per gene inside the call (outward neighbour scan with early
 * exit instead of all n^2 distances): an independent second implementation used as the cross-check of
 * gc_mi_cc_pairs (which measures faster on GeneClust's PCA embeddings, where a few isolated samples make one lane
 * of every warp walk the whole array).
 * scratch: gc_mi_cc_pairs_scratch_bytes(N, n), 256-byte aligned. */
int64_t gc_mi_cc_pairs_scratch_bytes(int64_t N, int n);
int gc_mi_cc_pairs_sorted(const double *XR, const double *YR, int64_t N, int n, int64_t ldr, const double *psi,
                          int64_t p_begin, int64_t p_end, double *mi_condensed, double *dense, uint8_t *nx_out,
                          uint8_t *ny_out, void *scratch, void *stream);
/* KSG MI for arbitrary-length problems (complementarity: one problem per (MST edge, cell cluster)):
 * problem q uses x = XR + x_off[q], y = YR + y_off[q], length len[q].  nx/ny scratch: int32, sum(len).
 * out_off[q] = offset of problem q in nx/ny.  mi[q] written. */
int gc_mi_cc_problems(const double *XR, const double *YR, const int64_t *x_off, const int64_t *y_off,
                      const int32_t *len, const int64_t *out_off, int64_t n_prob, const double *psi,
                      int32_t *nx, int32_t *ny, double *mi, void *stream);
/* Ross MI of G prepared feature vectors (XR[g*ldr + s], s < n) against integer labels.
 * label[s] in [0, n_label); label_k[l] = min(3, count_l - 1) (0 => samples of that label are ignored, they
 * are the "unique label" points).  const_term = digamma(n_kept) + mean digamma(k) - mean digamma(count)
 * (host, gene independent).  m_all: int32 scratch, n x G ([s*G+g]); mi[g] written. */
int gc_mi_cd(const double *XR, int64_t ldr, int64_t G, int n, const int32_t *label, const int32_t *label_k,
             int n_label, double const_term, const double *psi, int32_t *m_all, double *mi, void *stream);

/* Maximum-redundancy spanning forest of the graph with an edge wherever R != 0 (R: dense symmetric N x N float64 on
 * the device) - igraph Weighted_Adjacency(-R).spanning_tree at scGeneClust/tl/information.py:124-126.  Boruvka under
 * the strict edge order (larger R first, ties by smaller i*N + j, i < j), so the forest is unique.  edges_out:
 * (N - 1) x 2 int32 (unordered; sort by (i, j) for igraph's edge order), n_edges_out[0] = number of edges found
 * (N - 1 iff the graph is connected).  scratch: gc_mst_scratch_bytes(N), 256-byte aligned.  No host synchronisation. */
int64_t gc_mst_scratch_bytes(int64_t N);
int gc_mst_boruvka(const double *R, int64_t N, int32_t *edges_out, int32_t *n_edges_out, void *scratch, void *stream);
/* The same forest from edge lists, so that the N x N matrix never has to exist: the edges are the condensed pair range
 * [p_begin, p_begin + n_cond) (itertools.combinations order, values cond[0..n_cond)) plus n_extra explicit edges
 * (extra_u, extra_v, extra_w) - the forest carried over from earlier tiles or gathered from other ranks.  Non-positive
 * weights are non-edges.  Same strict edge order as gc_mst_boruvka (larger MI, then smaller condensed index), hence the
 * same forest: MSF(E1 u E2) = MSF(MSF(E1) u E2).  Outputs up to N - 1 edges (u < v) with their weights; *n_out = count.
 * The output arrays must not alias the explicit input edges.  scratch: gc_mst_scratch_bytes(N). */
int gc_mst_boruvka_edges(const double *cond, int64_t p_begin, int64_t n_cond, const int32_t *extra_u,
                         const int32_t *extra_v, const double *extra_w, int64_t n_extra, int64_t N, int32_t *out_u,
                         int32_t *out_v, double *out_w, int32_t *n_out, void *scratch, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * f2  High-confidence cells (scGeneClust/tl/confidence.py:58-60, 104-115): the cells x cells co-membership frequency
 * matrix of R Gaussian-mixture fits.  labels: R x C int32 (mixture component of every cell per fit), probas: R x C
 * float64 (its maximum posterior).  freq (C x C uint8, row-major): for j > i the number of fits in which
 * proba_i >= p, proba_j > p and label_i == label_j; zero elsewhere.  scratch: 2 R C int32.  R <= 32.
 * ------------------------------------------------------------------------------------------------------- */
int gc_comembership_counts(const int32_t *labels, const double *probas, int64_t C, int R, double p, uint8_t *freq,
                           int32_t *scratch, void *stream);

/* HOST function (all pointers are host memory, no GPU work): the HDBSCAN tail of GeneClust-ps
 * (scGeneClust/tl/cluster.py:127-130: hdbscan label -> condense_tree(., min_cluster_size) -> compute_stability ->
 * get_clusters('eom', allow_single_cluster=False) labelling).  mst_sorted_host: n_edges x 3 (i, j, weight), ascending by
 * weight, n_edges + 1 points.  Outputs: the condensed tree (parent/child/lambda/size, capacity 2 * n_edges rows,
 * n_rows_host[0] rows written - the input of the GLOSH scores) and labels_host[n_edges + 1] (-1 = noise). */
int gc_host_hdbscan_tree(const double *mst_sorted_host, int64_t n_edges, int64_t min_cluster_size, int64_t *parent_host,
                         int64_t *child_host, double *lambda_host, int64_t *size_host, int64_t *n_rows_host,
                         int64_t *labels_host);

/* HOST function: summed isolation depths of sklearn IsolationForest(random_state=seed).fit(x.reshape(-1, 1)) for
 * n <= 256 one-dimensional float32 samples - the singleton-rescue outlier test of scGeneClust/tl/selection.py:51
 * (numpy MT19937 seeding chain, sklearn's xorshift rand_r, depth-first random splits; see tl/selection.py).
 * c_of_host[m] = _average_path_length(m), m = 0..n (from the caller, so its log is numpy's); depths_host[n]. */
int gc_host_iforest_depths(const float *x_host, int n, uint32_t seed, int n_estimators, const double *c_of_host,
                           double *depths_host);

/* ---------------------------------------------------------------------------------------------------------
 * Synthetic negative-binomial counts (bench/test harness): counter-based, bit-identical
 * to oracle/synth.py.  Writes rows [c0, c0+nc) x genes [g0, g0+ng) of the global matrix into counts[nc x ld].
 * ------------------------------------------------------------------------------------------------------- */
int gc_synth_counts(uint64_t seed, int64_t c0, int64_t nc, int64_t g0, int64_t ng, const double *gene_mean,
                    const double *type_fc /* G_total x n_types */, int n_types, const int32_t *gene_module,
                    const double *size_levels, int n_size, const double *module_levels, int n_modlev,
                    int n_modules, uint16_t *counts, int64_t ld, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GENECLUST_B200_H */
