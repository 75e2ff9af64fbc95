// numpy's pairwise summation on the device, shared by the kernels that must round like a numpy reduction.
#pragma once
#include "common.cuh"

namespace gc {

// ---------------------------------------------------------------------------------------------------------
// numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum), bit-exact.
// `get(i)` returns element i of the (virtual) contiguous array.
// ---------------------------------------------------------------------------------------------------------
template <typename T, typename F>
__device__ __forceinline__ T np_pairwise_leaf(const F &get, int64_t lo, int64_t n) {
    if (n < 8) {
        T res = T(0);
        for (int64_t i = 0; i < n; ++i) res = res + get(lo + i);
        return res;
    }
    T r0 = get(lo + 0), r1 = get(lo + 1), r2 = get(lo + 2), r3 = get(lo + 3);
    T r4 = get(lo + 4), r5 = get(lo + 5), r6 = get(lo + 6), r7 = get(lo + 7);
    int64_t i = 8;
    const int64_t lim = n - (n % 8);
    for (; i < lim; i += 8) {
        r0 = r0 + get(lo + i + 0); r1 = r1 + get(lo + i + 1);
        r2 = r2 + get(lo + i + 2); r3 = r3 + get(lo + i + 3);
        r4 = r4 + get(lo + i + 4); r5 = r5 + get(lo + i + 5);
        r6 = r6 + get(lo + i + 6); r7 = r7 + get(lo + i + 7);
    }
    T res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res = res + get(lo + i);
    return res;
}

// The recursion of numpy (split at n/2 rounded down to a multiple of 8 until n <= 128) as an explicit-stack
// post-order walk: device recursion would need a stack the compiler cannot size.
template <typename T, typename F>
__device__ T np_pairwise_sum(const F &get, int64_t lo, int64_t n) {
    if (n <= 128) return np_pairwise_leaf<T>(get, lo, n);
    constexpr int kDepth = 40;
    int64_t s_lo[kDepth], s_n[kDepth];
    T s_left[kDepth];
    unsigned char s_state[kDepth];   // 1 = left child in progress, 2 = right child in progress
    int sp = 0;
    s_lo[0] = lo; s_n[0] = n;
    T result = T(0);
    while (true) {
        while (s_n[sp] > 128) {      // descend to the leftmost leaf
            int64_t n2 = s_n[sp] / 2;
            n2 -= n2 % 8;
            s_state[sp] = 1;
            s_lo[sp + 1] = s_lo[sp]; s_n[sp + 1] = n2;
            ++sp;
        }
        result = np_pairwise_leaf<T>(get, s_lo[sp], s_n[sp]);
        --sp;
        while (sp >= 0 && s_state[sp] == 2) { result = s_left[sp] + result; --sp; }
        if (sp < 0) break;
        s_left[sp] = result;         // left subtree done: walk the right one
        s_state[sp] = 2;
        int64_t n2 = s_n[sp] / 2;
        n2 -= n2 % 8;
        s_lo[sp + 1] = s_lo[sp] + n2; s_n[sp + 1] = s_n[sp] - n2;
        ++sp;
    }
    return result;
}

}  // namespace gc
