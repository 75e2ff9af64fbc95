// Maximum-redundancy spanning forest of the gene graph on the device (GeneClust-ps):
//   G = ig.Graph.Weighted_Adjacency(-R, mode="undirected"); MST = G.spanning_tree(weights=G.es["weight"])
// (scGeneClust/tl/information.py:124-126).  R is the dense symmetric N x N float64 redundancy matrix that already
// lives on the device; an edge exists only where R != 0 (MI is floored at 0, so zero entries are non-edges and the
// result is a spanning FOREST when that graph is disconnected).
//
// Boruvka with a strict total order on edges - larger R first, ties by smaller edge id i*N + j (i < j), the row-major
// upper-triangle order igraph numbers its edges in - so the forest is unique and independent of the execution order.
// ceil(log2 N) + 1 rounds, each: (1) one warp per vertex scans its row for the best edge leaving its component,
// (2) per-component best by two 64-bit atomics (weight bits, then edge id among the weight ties), (3) the component
// roots emit their edge (a mutually chosen edge once) and hook, (4) pointer jumping.  No host round trip.
#include "common.cuh"

#include <algorithm>

namespace gc {

struct MstState {
    int32_t *comp;                  // component root of every vertex
    unsigned long long *cand_w;     // per vertex: bits of the best R leaving its component (0 = none)
    unsigned long long *cand_e;     // per vertex: its edge id
    unsigned long long *best_w;     // per component root
    unsigned long long *best_e;
    int32_t *hook;                  // per component root: the root it merges into (itself = stays)
};

__global__ void mst_init_kernel(MstState s, int N, int32_t *n_edges) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < N) s.comp[v] = v;
    if (v == 0) *n_edges = 0;
}

__global__ void __launch_bounds__(256)
mst_vertex_best_kernel(const double *__restrict__ R, int N, MstState s) {
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= N) return;
    const int cv = s.comp[v];
    if (lane == 0 && cv == v) { s.best_w[v] = 0ull; s.best_e[v] = ~0ull; }     // reset for this round (roots only)
    unsigned long long bw = 0ull, be = ~0ull;
    const double *row = R + (int64_t)v * N;
    for (int u = lane; u < N; u += 32) {
        const double r = row[u];
        if (r > 0.0 && s.comp[u] != cv) {
            const unsigned long long w = (unsigned long long)__double_as_longlong(r);     // monotone for r > 0
            const unsigned long long e = v < u ? (unsigned long long)v * N + u : (unsigned long long)u * N + v;
            if (w > bw || (w == bw && e < be)) { bw = w; be = e; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ow = __shfl_xor_sync(0xffffffffu, bw, o), oe = __shfl_xor_sync(0xffffffffu, be, o);
        if (ow > bw || (ow == bw && oe < be)) { bw = ow; be = oe; }
    }
    if (lane == 0) { s.cand_w[v] = bw; s.cand_e[v] = be; }
}

__global__ void mst_comp_weight_kernel(int N, MstState s) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < N && s.cand_w[v] != 0ull) atomicMax(&s.best_w[s.comp[v]], s.cand_w[v]);
}
__global__ void mst_comp_edge_kernel(int N, MstState s) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < N && s.cand_w[v] != 0ull && s.cand_w[v] == s.best_w[s.comp[v]]) atomicMin(&s.best_e[s.comp[v]], s.cand_e[v]);
}

// roots: emit the chosen edge (a mutually chosen edge only from the smaller root) and decide the hook
__global__ void mst_hook_kernel(int N, MstState s, int32_t *__restrict__ edges, int32_t *n_edges) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    if (s.comp[c] != c) return;
    s.hook[c] = c;
    if (s.best_w[c] == 0ull) return;                     // isolated component: no edge leaves it
    const unsigned long long e = s.best_e[c];
    const int a = (int)(e / (unsigned long long)N), b = (int)(e % (unsigned long long)N);
    const int ca = s.comp[a], cb = s.comp[b];
    const int other = ca == c ? cb : ca;
    const bool mutual = s.best_w[other] != 0ull && s.best_e[other] == e;
    if (mutual && c < other) return;                     // the larger root emits the shared edge and hooks to us
    const int idx = atomicAdd(n_edges, 1);
    edges[2 * idx] = a;
    edges[2 * idx + 1] = b;
    s.hook[c] = other;
}

// comp[v] <- root of the hook forest (chains have length <= N; every step at least halves them)
__global__ void mst_jump_kernel(int N, MstState s, int32_t *changed) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N) return;
    int c = s.comp[v];
    int r = s.hook[c];
    while (s.hook[r] != r) r = s.hook[r];                // hook[] is only written by mst_hook_kernel: a static forest here
    if (r != c) { s.comp[v] = r; if (changed) *changed = 1; }
}


// ---------------------------------------------------------------------------------------------------------
// The same forest from EDGE LISTS, so that the dense N x N matrix never has to exist (north_star: "no dense N x N
// materialisation").  The edges of one call are
//   * a contiguous range [p_begin, p_begin + n_cond) of the condensed pair index (itertools.combinations order, the
//     order the redundancy kernel writes its values in) with their MI values, and
//   * an explicit list (u, v, w) of carried edges - the forest found so far.
// Because MSF(E1 u E2) = MSF(MSF(E1) u E2) under a strict total edge order, the redundancy stage can stream tiles of the
// pair space through this routine (single GPU), and ranks that own disjoint pair ranges can each reduce their range to
// <= N - 1 edges, all-gather those and reduce once more (the "top-k merge" of north_star) - every rank ends with the
// forest the dense routine above finds.  Edge order: larger MI first, ties by smaller condensed index (= the row-major
// upper-triangle order of the dense routine, so both give the same forest).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mst_pair_from_index(long long p, long long N, int &i, int &j) {
    const double b = 2.0 * (double)N - 1.0;
    long long ii = (long long)floor((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
    if (ii < 0) ii = 0;
    if (ii > N - 2) ii = N - 2;
    while (ii > 0 && ii * N - ii * (ii + 1) / 2 > p) --ii;
    while ((ii + 1) * N - (ii + 1) * (ii + 2) / 2 <= p) ++ii;
    i = (int)ii;
    j = (int)(p - (ii * N - ii * (ii + 1) / 2) + ii + 1);
}
__device__ __forceinline__ long long mst_index_from_pair(long long u, long long v, long long N) {
    const long long a = u < v ? u : v, b = u < v ? v : u;
    return a * N - a * (a + 1) / 2 + (b - a - 1);
}

struct MstEdges {
    const double *cond; long long p_begin, n_cond;           // implicit edges
    const int32_t *eu, *ev; const double *ew; long long n_extra;   // explicit edges
};
// edge t of the union: endpoints, weight, canonical id
__device__ __forceinline__ bool mst_edge(const MstEdges &E, long long t, long long N, int &u, int &v, double &w, long long &id) {
    if (t < E.n_cond) {
        w = E.cond[t];
        if (!(w > 0.0)) return false;                        // MI is floored at 0: zero entries are non-edges
        id = E.p_begin + t;
        mst_pair_from_index(id, N, u, v);
        return true;
    }
    const long long x = t - E.n_cond;
    w = E.ew[x];
    if (!(w > 0.0)) return false;
    u = E.eu[x]; v = E.ev[x];
    id = mst_index_from_pair(u, v, N);
    return true;
}

__global__ void mst_reset_roots_kernel(int N, MstState s) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < N && s.comp[v] == v) { s.best_w[v] = 0ull; s.best_e[v] = ~0ull; }
}
__global__ void __launch_bounds__(256) mst_edges_weight_kernel(MstEdges E, long long n_edges, int N, MstState s) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_edges; t += stride) {
        int u, v; double w; long long id;
        if (!mst_edge(E, t, N, u, v, w, id)) continue;
        const int cu = s.comp[u], cv = s.comp[v];
        if (cu == cv) continue;
        const unsigned long long wb = (unsigned long long)__double_as_longlong(w);
        if (wb > s.best_w[cu]) atomicMax(&s.best_w[cu], wb);
        if (wb > s.best_w[cv]) atomicMax(&s.best_w[cv], wb);
    }
}
__global__ void __launch_bounds__(256) mst_edges_id_kernel(MstEdges E, long long n_edges, int N, MstState s) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_edges; t += stride) {
        int u, v; double w; long long id;
        if (!mst_edge(E, t, N, u, v, w, id)) continue;
        const int cu = s.comp[u], cv = s.comp[v];
        if (cu == cv) continue;
        const unsigned long long wb = (unsigned long long)__double_as_longlong(w);
        if (wb == s.best_w[cu]) atomicMin(&s.best_e[cu], (unsigned long long)id);
        if (wb == s.best_w[cv]) atomicMin(&s.best_e[cv], (unsigned long long)id);
    }
}
// roots: emit the chosen edge with its weight (a mutually chosen edge only once) and decide the hook
__global__ void mst_hook_edges_kernel(int N, MstState s, int32_t *__restrict__ out_u, int32_t *__restrict__ out_v,
                                      double *__restrict__ out_w, int32_t *n_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    if (s.comp[c] != c) return;
    s.hook[c] = c;
    if (s.best_w[c] == 0ull) return;
    const unsigned long long e = s.best_e[c];
    int a, b;
    mst_pair_from_index((long long)e, N, a, b);
    const int ca = s.comp[a], cb = s.comp[b];
    const int other = ca == c ? cb : ca;
    const bool mutual = s.best_w[other] != 0ull && s.best_e[other] == e;
    if (mutual && c < other) return;
    const int idx = atomicAdd(n_out, 1);
    out_u[idx] = a; out_v[idx] = b;
    out_w[idx] = __longlong_as_double((long long)s.best_w[c]);
    s.hook[c] = other;
}

}  // namespace gc

using namespace gc;

static int64_t mst_align(int64_t x) { return (x + 255) / 256 * 256; }

extern "C" int64_t gc_mst_scratch_bytes(int64_t N) {
    return 2 * mst_align(N * 4) + 4 * mst_align(N * 8) + 512;
}

extern "C" int gc_mst_boruvka(const double *R, int64_t N, int32_t *edges_out, int32_t *n_edges_out, void *scratch,
                              void *stream) {
    GC_REQUIRE(R && edges_out && n_edges_out && scratch, "null pointer");
    GC_REQUIRE(N >= 1 && N < (1ll << 30), "bad shape");
    GC_REQUIRE(((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    char *p = (char *)scratch;
    MstState s;
    s.comp = (int32_t *)p;                    p += mst_align(N * 4);
    s.hook = (int32_t *)p;                    p += mst_align(N * 4);
    s.cand_w = (unsigned long long *)p;       p += mst_align(N * 8);
    s.cand_e = (unsigned long long *)p;       p += mst_align(N * 8);
    s.best_w = (unsigned long long *)p;       p += mst_align(N * 8);
    s.best_e = (unsigned long long *)p;
    const unsigned nb = (unsigned)ceil_div<int64_t>(N, 256);
    mst_init_kernel<<<nb, 256, 0, st>>>(s, (int)N, n_edges_out);
    if (int rc = check_launch("mst_init_kernel")) return rc;
    int rounds = 1;
    while ((1ll << rounds) < N) ++rounds;
    ++rounds;                                            // every round at least halves the number of mergeable components
    for (int r = 0; r < rounds; ++r) {
        mst_vertex_best_kernel<<<(unsigned)ceil_div<int64_t>(N * 32, 256), 256, 0, st>>>(R, (int)N, s);
        if (int rc = check_launch("mst_vertex_best_kernel")) return rc;
        mst_comp_weight_kernel<<<nb, 256, 0, st>>>((int)N, s);
        if (int rc = check_launch("mst_comp_weight_kernel")) return rc;
        mst_comp_edge_kernel<<<nb, 256, 0, st>>>((int)N, s);
        if (int rc = check_launch("mst_comp_edge_kernel")) return rc;
        mst_hook_kernel<<<nb, 256, 0, st>>>((int)N, s, edges_out, n_edges_out);
        if (int rc = check_launch("mst_hook_kernel")) return rc;
        mst_jump_kernel<<<nb, 256, 0, st>>>((int)N, s, nullptr);
        if (int rc = check_launch("mst_jump_kernel")) return rc;
    }
    return 0;
}

extern "C" int gc_mst_boruvka_edges(const double *cond, int64_t p_begin, int64_t n_cond, const int32_t *extra_u,
                                    const int32_t *extra_v, const double *extra_w, int64_t n_extra, int64_t N,
                                    int32_t *out_u, int32_t *out_v, double *out_w, int32_t *n_out, void *scratch,
                                    void *stream) {
    GC_REQUIRE(out_u && out_v && out_w && n_out && scratch, "null pointer");
    GC_REQUIRE(N >= 1 && N < (1ll << 30) && n_cond >= 0 && n_extra >= 0 && p_begin >= 0, "bad shape");
    GC_REQUIRE(n_cond == 0 || cond, "null condensed values");
    GC_REQUIRE(n_extra == 0 || (extra_u && extra_v && extra_w), "null explicit edges");
    GC_REQUIRE(p_begin + n_cond <= N * (N - 1) / 2, "pair range outside the condensed index space");
    GC_REQUIRE(((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    char *p = (char *)scratch;
    MstState s;
    s.comp = (int32_t *)p;                    p += mst_align(N * 4);
    s.hook = (int32_t *)p;                    p += mst_align(N * 4);
    s.cand_w = (unsigned long long *)p;       p += mst_align(N * 8);
    s.cand_e = (unsigned long long *)p;       p += mst_align(N * 8);
    s.best_w = (unsigned long long *)p;       p += mst_align(N * 8);
    s.best_e = (unsigned long long *)p;
    MstEdges E{cond, (long long)p_begin, (long long)n_cond, extra_u, extra_v, extra_w, (long long)n_extra};
    const long long n_edges = (long long)n_cond + (long long)n_extra;
    const unsigned nb = (unsigned)ceil_div<int64_t>(N, 256);
    const unsigned eb = (unsigned)std::max<long long>(1, std::min<long long>(ceil_div<long long>(n_edges, 256), (long long)kNumSMs * 16));
    mst_init_kernel<<<nb, 256, 0, st>>>(s, (int)N, n_out);
    if (int rc = check_launch("mst_init_kernel")) return rc;
    if (n_edges == 0) return 0;
    int rounds = 1;
    while ((1ll << rounds) < N) ++rounds;
    ++rounds;
    for (int r = 0; r < rounds; ++r) {
        mst_reset_roots_kernel<<<nb, 256, 0, st>>>((int)N, s);
        if (int rc = check_launch("mst_reset_roots_kernel")) return rc;
        mst_edges_weight_kernel<<<eb, 256, 0, st>>>(E, n_edges, (int)N, s);
        if (int rc = check_launch("mst_edges_weight_kernel")) return rc;
        mst_edges_id_kernel<<<eb, 256, 0, st>>>(E, n_edges, (int)N, s);
        if (int rc = check_launch("mst_edges_id_kernel")) return rc;
        mst_hook_edges_kernel<<<nb, 256, 0, st>>>((int)N, s, out_u, out_v, out_w, n_out);
        if (int rc = check_launch("mst_hook_edges_kernel")) return rc;
        mst_jump_kernel<<<nb, 256, 0, st>>>((int)N, s, nullptr);
        if (int rc = check_launch("mst_jump_kernel")) return rc;
    }
    return 0;
}
