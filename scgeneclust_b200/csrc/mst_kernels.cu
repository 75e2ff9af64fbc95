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
