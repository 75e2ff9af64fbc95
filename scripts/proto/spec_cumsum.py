"""CPU prototype of the speculative exact float32 cumulative sum used by the k-means++ round kernel (developer tool).

Sequential float32 `np.cumsum` is a dependent chain.  Pieces of the array are summed concurrently from SURROGATE start
values (the bottom of the guessed binade, with either parity of the last mantissa bit); inside one binade the
increment a piece adds depends only on that parity, so  end = start + D[parity]  - checked afterwards (start really in
the guessed binade, end still in it), and a piece that fails the check is summed for real.  This script checks the
result against np.cumsum bit for bit on random and adversarial inputs."""
import sys

import numpy as np

f32 = np.float32


def expo(x):
    return (np.asarray(x, dtype=f32).view(np.uint32) >> 23) & 0xFF


def spec_cumsum_ends(v, carry, P):
    """Returns (starts[P], ends[P], n_real): exact sequential float32 running sums at the piece boundaries."""
    cn = len(v)
    ln = ((cn + P - 1) // P) | 1
    pad = np.zeros(P * ln, dtype=f32)
    pad[:cn] = v
    pieces = pad.reshape(P, ln)
    approx = np.concatenate([[0.0], np.cumsum(pieces.sum(1, dtype=np.float64))[:-1]]) + float(carry)
    e = expo(approx.astype(f32)).astype(np.int64)
    ok_guess = (e >= 1) & (e <= 253)
    a0 = (e.astype(np.uint32) << 23).view(f32)
    b0 = ((e.astype(np.uint32) << 23) + 1).view(f32)
    a, b = a0.copy(), b0.copy()
    for i in range(ln):
        a = (a + pieces[:, i]).astype(f32)
        b = (b + pieces[:, i]).astype(f32)
    with np.errstate(invalid='ignore', over='ignore'):
        d0, d1 = (a - a0).astype(f32), (b - b0).astype(f32)
    starts, ends = np.zeros(P, dtype=f32), np.zeros(P, dtype=f32)
    n_real = 0

    def real(start, p):
        acc = f32(start)
        for x in pieces[p]:
            acc = f32(acc + x)
        return acc
    start = f32(carry)
    for p in range(P):
        starts[p] = start
        sb = int(np.asarray(start, dtype=f32).view(np.uint32))
        es = (sb >> 23) & 0xFF
        d = d1[p] if (sb & 1) else d0[p]
        with np.errstate(invalid='ignore', over='ignore'):
            end = f32(start + d)
        ok = p > 0 and ok_guess[p] and es == e[p] and int(expo(end)) == es
        if not ok:
            end = real(start, p)
            n_real += 1
        ends[p] = end
        start = end
    return starts, ends, n_real, ln


def check(v, carry=0.0, P=64):
    v = np.asarray(v, dtype=f32)
    seq = np.cumsum(np.concatenate([[f32(carry)], v]).astype(f32), dtype=f32)[1:]     # sequential float32 chain from carry
    # np.cumsum with float32 accumulates sequentially (pairwise summation is only used by add.reduce)
    acc, ref = f32(carry), np.empty(len(v), dtype=f32)
    for i, x in enumerate(v[:2000]):
        acc = f32(acc + x)
        ref[i] = acc
    assert np.array_equal(ref[:min(len(v), 2000)].view(np.uint32), seq[:min(len(v), 2000)].view(np.uint32))
    starts, ends, n_real, ln = spec_cumsum_ends(v, carry, P)
    for p in range(P):
        hi = min((p + 1) * ln, len(v))
        if hi <= 0:
            continue
        want = seq[hi - 1] if hi > p * ln or p == 0 else seq[-1]
        if p * ln >= len(v):
            want = seq[-1]
        assert ends[p].view(np.uint32) == want.view(np.uint32), (p, ends[p], want)
    return n_real


if __name__ == "__main__":
    rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    tot_real, n = 0, 0
    for trial in range(300):
        cn = int(rs.choice([6000, 6144, 5999, 3072, 100, 64, 1, 9000 % 6144]))
        kind = trial % 6
        if kind == 0:
            v = rs.gamma(2.0, 30.0, cn)                                # typical squared distances
        elif kind == 1:
            v = rs.gamma(2.0, 30.0, cn) * (rs.rand(cn) < 0.7)          # many exact zeros (chosen centres / duplicates)
        elif kind == 2:
            v = np.ldexp(rs.randint(1, 64, cn).astype(np.float64), rs.randint(-6, 3, cn))   # few mantissa bits: many ties
        elif kind == 3:
            v = rs.gamma(2.0, 30.0, cn); v[rs.randint(0, cn, 3)] *= 1e6                      # outliers >> running sum
        elif kind == 4:
            v = np.full(cn, 0.1) * rs.choice([1, 2, 4], cn)                                  # repeated values
        else:
            v = rs.lognormal(0, 4, cn)                                                       # 20 binades of dynamic range
        carry = 0.0 if trial % 2 == 0 else float(f32(rs.gamma(2.0, 1e5)))
        tot_real += check(v.astype(f32), carry)
        n += 1
    print(f"{n} cases bit-identical to the sequential float32 chain; pieces summed for real: {tot_real / n:.1f} of 64 on average")
