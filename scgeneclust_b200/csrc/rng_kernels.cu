// numpy.random.RandomState(seed).normal(size=(rows, q)).astype(float32) on the device, bit-for-bit the legacy stream:
// MT19937 -> rk_double pairs -> Marsaglia polar method with rejection (numpy/random/src/legacy/legacy-distributions.c
// legacy_gauss).  This is the Omega test matrix scikit-learn draws for its randomized SVD (utils/extmath.py:323-334),
// reached from sc.pp.pca at scGeneClust/pp/_preprocessing.py:52,54; on the host it costs 20-40 ms for 20 000 x 60
// (more than all sixteen passes over the count matrix), here ~1 ms.
//
// Why it parallelises although the generator is sequential: every polar ATTEMPT consumes exactly four 32-bit words
// whether it is accepted or not, so attempt i reads words [4i, 4i+4) and the accepted attempts, compacted in order,
// yield the normals (f*x2 first, then the cached f*x1).  Only the MT19937 word stream itself is sequential; one CTA
// produces it block by block (624 words per block, four dependent phases).
// Compiled with -fmad=false: r2 = x1*x1 + x2*x2 must round like the C expression (acceptance decisions are exact).
// log() may differ from glibc's in the last bit of the float64 result; after the float32 cast the panels are identical
// (checked against numpy in tests/test_gpu_embed.py).
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include "mt_jump_poly.h"

namespace gc {

constexpr int kMtN = 624, kMtM = 397;
constexpr int kAttemptsPerCta = 2048;    // 256 threads x 8 attempts

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// the twist term of the recurrence  x[k + 624] = x[k + 397] ^ tw(x[k], x[k + 1])
__device__ __forceinline__ uint32_t mt_tw(uint32_t cur, uint32_t nxt) {
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// One CTA: emits the n_words RAW (untempered) state words starting at position `pos` (0..624; 624 = regenerate first) of `state`.
// A 624-word block is regenerated from the PREVIOUS block alone, all 624 words in parallel, one __syncthreads per block
// (the textbook in-place loop needs three dependent phases: words [227, 454) read new words [0, 227), words [454, 624)
// read new words [227, 397)).  Substituting the recurrence into itself removes the dependence on new words:
//   new[i] = old[i + 397]                                                                 ^ tw(old[i], old[i + 1])   i < 227
//   new[i] = old[i + 170] ^ tw(old[i - 227], old[i - 226])                                ^ tw(old[i], old[i + 1])   227 <= i < 454
//   new[i] = old[i -  57] ^ tw(old[i - 454], old[i - 453]) ^ tw(old[i - 227], old[i - 226]) ^ tw(old[i], old[i + 1]) 454 <= i
// (for i = 623 the last argument is new[0] = old[397] ^ tw(old[0], old[1]), recomputed by that thread).  Checked word
// for word against numpy's stream (tests/test_gpu_embed.py compares the normals drawn from it).
constexpr int kMtThreads = 640;
__device__ __forceinline__ void mt_stream(const uint32_t *__restrict__ state, int pos, int64_t n_words,
                                          uint32_t *__restrict__ out, uint32_t (*buf)[kMtN]) {
    const int t = threadIdx.x;
    for (int i = t; i < kMtN; i += blockDim.x) buf[0][i] = state[i];
    __syncthreads();
    int64_t done = 0;
    if (pos < kMtN) {          // words left in the current block
        const int avail = kMtN - pos;
        for (int i = t; i < avail && i < n_words; i += blockDim.x) out[i] = buf[0][pos + i];       // RAW state words
        done = min((int64_t)avail, n_words);
    }
    int cur = 0;
    while (done < n_words) {
        const uint32_t *mt = buf[cur];
        uint32_t *nt = buf[cur ^ 1];
        const int64_t left = n_words - done;
        if (t < kMtN) {
            uint32_t v;
            if (t < 227) {
                v = mt[t + kMtM] ^ mt_tw(mt[t], mt[t + 1]);
            } else if (t < 454) {
                v = mt[t + 170] ^ mt_tw(mt[t - 227], mt[t - 226]) ^ mt_tw(mt[t], mt[t + 1]);
            } else {
                const uint32_t nxt = t < kMtN - 1 ? mt[t + 1] : (mt[kMtM] ^ mt_tw(mt[0], mt[1]));
                v = mt[t - 57] ^ mt_tw(mt[t - 454], mt[t - 453]) ^ mt_tw(mt[t - 227], mt[t - 226]) ^ mt_tw(mt[t], nxt);
            }
            nt[t] = v;
            if (t < left) out[done + t] = v;          // RAW: the (parallel) consumers temper on read
        }
        __syncthreads();
        done += min((int64_t)kMtN, left);
        cur ^= 1;
    }
}

__global__ void __launch_bounds__(kMtThreads)
mt19937_stream_kernel(const uint32_t *__restrict__ state, int pos, int64_t n_words, uint32_t *__restrict__ out) {
    __shared__ uint32_t buf[2][kMtN];
    mt_stream(state, pos, n_words, out, buf);
}

// ---- the stream in parallel segments ----------------------------------------------------------------------------
// The word sequence obeys a linear recurrence over GF(2), so the block that precedes word k + J is a fixed XOR
// combination of words k .. k + 19936 + 623 (jump polynomial x^J mod the characteristic polynomial, mt_jump_poly.h,
// generated and verified by scripts/mt_jump_poly.py).  With J = one segment (kMtSegBlocks blocks) less a block:
//   head(s):  the first kMtHeadBlocks = 33 blocks of segment s, from the block preceding it        (one CTA)
//   jump(s):  the block preceding segment s + 1, from head(s)                                        (39 CTAs)
// alternate down the segments (short kernels), then ONE launch lets every segment finish its remaining 479 blocks on
// its own CTA.  50 000 x 60 normals: 10 segments, ~0.3 ms instead of 1.36 ms on one CTA.
constexpr int kMtHeadBlocks = 33;                              // 33 * 624 = 20592 >= 19937 + 623
constexpr int64_t kMtSegWords = (int64_t)kMtSegBlocks * kMtN;

// segments [first, first + gridDim.x): the words after the head
__global__ void __launch_bounds__(kMtThreads)
mt19937_body_kernel(uint32_t *__restrict__ words, int64_t n_words) {
    __shared__ uint32_t buf[2][kMtN];
    const int64_t seg0 = (int64_t)blockIdx.x * kMtSegWords;
    const int64_t seg_len = min(kMtSegWords, n_words - seg0);
    const int64_t head = (int64_t)kMtHeadBlocks * kMtN;
    if (seg_len <= head) return;
    mt_stream(words + seg0 + head - kMtN, kMtN, seg_len - head, words + seg0 + head, buf);
}

// prev[w] = XOR over the set coefficients i of the jump polynomial of y[i + w],  w = 0..623
__global__ void __launch_bounds__(1024)
mt19937_jump_kernel(const uint32_t *__restrict__ y, uint32_t *__restrict__ prev) {
    __shared__ uint32_t part[64][16];
    const int wi = threadIdx.x & 15, chunk = threadIdx.x >> 4;
    const int w = blockIdx.x * 16 + wi;                          // 39 CTAs x 16 words
    // 624 polynomial words over 64 chunks: ten words (320 coefficients) per thread, the last chunk takes what is left;
    // the 32 loads of a word are independent (masked, not predicated), so they are all in flight together
    constexpr int kWordsPer = 10;
    const int pw_lo = chunk * kWordsPer, pw_hi = min(kMtN, pw_lo + kWordsPer);
    uint32_t acc = 0;
    if (w < kMtN) {
        for (int pw = pw_lo; pw < pw_hi; ++pw) {
            const uint32_t g = kMtJumpPoly[pw];                    // coefficients 32 pw .. 32 pw + 31 (zero above 19936)
            const uint32_t *yy = y + 32 * pw + w;
#pragma unroll
            for (int b = 0; b < 32; ++b) acc ^= yy[b] & (0u - ((g >> b) & 1u));
        }
    }
    part[chunk][wi] = acc;
    __syncthreads();
    if (threadIdx.x < 16 && blockIdx.x * 16 + (int)threadIdx.x < kMtN) {
        uint32_t v = 0;
#pragma unroll 8
        for (int c = 0; c < 64; ++c) v ^= part[c][threadIdx.x];
        prev[blockIdx.x * 16 + threadIdx.x] = v;
    }
}

__device__ __forceinline__ double rk_double(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}
// attempt i: x1, x2 in (-1, 1), accepted iff 0 < r2 < 1
__device__ __forceinline__ bool polar_attempt(const uint32_t *__restrict__ w, int64_t i, double &x1, double &x2, double &r2) {
    // the stream kernel stores the raw MT19937 state words (it is one CTA: every instruction it does not execute shortens
    // the sequential part); the tempering of the output function happens here, in the massively parallel consumers
    const uint4 q = reinterpret_cast<const uint4 *>(w)[i];
    x1 = 2.0 * rk_double(mt_temper(q.x), mt_temper(q.y)) - 1.0;
    x2 = 2.0 * rk_double(mt_temper(q.z), mt_temper(q.w)) - 1.0;
    r2 = x1 * x1 + x2 * x2;
    return r2 < 1.0 && r2 != 0.0;
}

__global__ void __launch_bounds__(256)
polar_count_kernel(const uint32_t *__restrict__ words, int64_t n_attempts, int32_t *__restrict__ cta_count) {
    __shared__ int warp_cnt[8];
    const int64_t base = (int64_t)blockIdx.x * kAttemptsPerCta;
    int c = 0;
    for (int k = 0; k < kAttemptsPerCta / 256; ++k) {
        const int64_t i = base + k * 256 + threadIdx.x;
        double x1, x2, r2;
        if (i < n_attempts && polar_attempt(words, i, x1, x2, r2)) ++c;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += warp_cnt[w];
        cta_count[blockIdx.x] = s;
    }
}

// exclusive scan of the per-CTA counts (single CTA, sequential over <= a few thousand entries per thread chunk)
__global__ void __launch_bounds__(1024)
polar_scan_kernel(const int32_t *__restrict__ cta_count, int n_cta, int64_t *__restrict__ cta_offset, int64_t n_pairs_needed,
                  int32_t *__restrict__ status) {
    __shared__ int64_t part[1024];
    const int per = (n_cta + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(n_cta, lo + per);
    int64_t s = 0;
    for (int i = lo; i < hi; ++i) s += cta_count[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int i = 0; i < 1024; ++i) { const int64_t v = part[i]; part[i] = run; run += v; }
        if (run < n_pairs_needed) status[0] |= 2;      // not enough accepted attempts in the generated stream
    }
    __syncthreads();
    int64_t run = part[threadIdx.x];
    for (int i = lo; i < hi; ++i) { cta_offset[i] = run; run += cta_count[i]; }
}

// accepted attempt number p (0-based, stream order) yields normals 2p (= f*x2) and 2p+1 (= f*x1); normal m goes to
// panel[(m / q) * ld + m % q]
__global__ void __launch_bounds__(256)
polar_emit_kernel(const uint32_t *__restrict__ words, int64_t n_attempts, const int64_t *__restrict__ cta_offset,
                  int64_t n_normals, int q, int ld, float *__restrict__ panel) {
    __shared__ int warp_cnt[8];
    __shared__ int64_t s_base;
    const int64_t base = (int64_t)blockIdx.x * kAttemptsPerCta;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = cta_offset[blockIdx.x];
    int64_t run = 0;                           // accepted so far in this CTA (before the current 256-attempt slab)
    for (int k = 0; k < kAttemptsPerCta / 256; ++k) {
        const int64_t i = base + k * 256 + threadIdx.x;
        double x1 = 0, x2 = 0, r2 = 1;
        const bool ok = i < n_attempts && polar_attempt(words, i, x1, x2, r2);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 8; ++w) { const int c = warp_cnt[w]; if (w < warp) before += c; total += c; }
        if (ok) {
            const int64_t p = s_base + run + before + __popc(m & ((1u << lane) - 1u));
            const double f = sqrt(-2.0 * log(r2) / r2);
            const int64_t m0 = 2 * p, m1 = 2 * p + 1;
            if (m0 < n_normals) panel[(m0 / q) * ld + (m0 % q)] = (float)(f * x2);
            if (m1 < n_normals) panel[(m1 / q) * ld + (m1 % q)] = (float)(f * x1);
        }
        run += total;
        __syncthreads();
    }
}

}  // namespace gc

using namespace gc;

static int64_t rng_attempts(int64_t n_normals) {
    // acceptance probability pi/4; margin of 1 % + 8192 attempts (the count's standard deviation is ~0.37 sqrt(n))
    const double need = (double)((n_normals + 1) / 2) * (4.0 / 3.14159265358979323846);
    return (int64_t)(need * 1.01) + 8192;
}
static int64_t rng_align(int64_t x) { return (x + 255) / 256 * 256; }

extern "C" int64_t gc_mt19937_normal_scratch_bytes(int64_t rows, int q) {
    const int64_t att = rng_attempts(rows * q);
    const int64_t n_cta = ceil_div<int64_t>(att, kAttemptsPerCta);
    return rng_align(att * 16) + rng_align(n_cta * 4) + rng_align(n_cta * 8) + rng_align(kMtN * 4) + 256;
}

// n_words raw MT19937 state words following position `pos` of `state`; prev: 624 words of scratch
static int mt19937_words(const uint32_t *state, int pos, int64_t n_words, uint32_t *words, uint32_t *prev, cudaStream_t st) {
    static const bool serial = getenv("GC_MT_SERIAL") != nullptr;          // developer knob: the one-CTA stream
    const int64_t n_seg = ceil_div<int64_t>(n_words, kMtSegWords);
    if (pos != kMtN || n_seg < 2 || serial) {
        mt19937_stream_kernel<<<1, kMtThreads, 0, st>>>(state, pos, n_words, words);
        return check_launch("mt19937_stream_kernel");
    }
    const int64_t head = (int64_t)kMtHeadBlocks * kMtN;
    for (int64_t s = 0; s < n_seg; ++s) {
        const int64_t seg0 = s * kMtSegWords;
        mt19937_stream_kernel<<<1, kMtThreads, 0, st>>>(s == 0 ? state : prev, kMtN, std::min(head, n_words - seg0), words + seg0);
        if (int rc = check_launch("mt19937_stream_kernel")) return rc;
        if (s + 1 < n_seg) {
            mt19937_jump_kernel<<<(kMtN + 15) / 16, 1024, 0, st>>>(words + seg0, prev);
            if (int rc = check_launch("mt19937_jump_kernel")) return rc;
        }
    }
    mt19937_body_kernel<<<(unsigned)n_seg, kMtThreads, 0, st>>>(words, n_words);
    return check_launch("mt19937_body_kernel");
}

extern "C" int gc_mt19937_normal_panel(const uint32_t *mt_state, int mt_pos, int64_t rows, int q, float *panel, int ld,
                                       void *scratch, int32_t *status, void *stream) {
    GC_REQUIRE(mt_state && panel && scratch && status, "null pointer");
    GC_REQUIRE(mt_pos >= 0 && mt_pos <= kMtN, "mt_pos must be in [0, 624]");
    GC_REQUIRE(rows >= 1 && q >= 1 && ld >= q, "bad shape");
    GC_REQUIRE(((uintptr_t)scratch & 255) == 0, "scratch must be 256-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int64_t n_normals = rows * q, att = rng_attempts(n_normals);
    const int64_t n_cta = ceil_div<int64_t>(att, kAttemptsPerCta);
    GC_REQUIRE(n_cta < (1ll << 31), "panel too large");
    char *p = (char *)scratch;
    uint32_t *words = (uint32_t *)p;      p += rng_align(att * 16);
    int32_t *cta_count = (int32_t *)p;    p += rng_align(n_cta * 4);
    int64_t *cta_offset = (int64_t *)p;
    if (int rc = mt19937_words(mt_state, mt_pos, att * 4, words, (uint32_t *)(p + rng_align(n_cta * 8)), st)) return rc;
    polar_count_kernel<<<(unsigned)n_cta, 256, 0, st>>>(words, att, cta_count);
    if (int rc = check_launch("polar_count_kernel")) return rc;
    polar_scan_kernel<<<1, 1024, 0, st>>>(cta_count, (int)n_cta, cta_offset, (n_normals + 1) / 2, status);
    if (int rc = check_launch("polar_scan_kernel")) return rc;
    polar_emit_kernel<<<(unsigned)n_cta, 256, 0, st>>>(words, att, cta_offset, n_normals, q, ld, panel);
    return check_launch("polar_emit_kernel");
}
