"""Prototype of the piece-parallel, bit-exact sequential float32 sum used by the deviance column sums
(csrc/deviance_kernels.cu: colsum_spec_*).  numpy's add.reduce over the outer axis of a C-ordered (C, m) float32 array
adds row after row: acc = fl(acc + x_i).  While acc stays inside one binade [2^e, 2^(e+1)) (ulp u = 2^(e-23)) and keeps
its sign, fl(acc + x) = acc + sgn * rn_u(sgn * x) with integer arithmetic in units of u; only a tie (|x|/u has
fraction exactly 1/2) needs the parity of acc.  So every piece of rows can be summarised independently as two integer
deltas (acc even / odd on entry) with the extremes of the running offset, given a GUESS of (sign, e) for the piece -
taken from float64 prefix sums - and a sequential pass over the pieces verifies each guess against the true acc and
replays the piece add by add when it does not hold.  Exact by construction; the guess only decides the speed.

Run: python scripts/proto/spec_colsum.py   (compares with the sequential float32 loop on adversarial inputs)"""
import numpy as np


def seq_sum(x):
    acc = np.float32(x[0])
    for v in x[1:]:
        acc = np.float32(acc + v)
    return acc


def decompose(v):
    """float32 -> (sign, M, ex): |v| = M * 2^ex, M < 2^24."""
    b = int(np.float32(v).view(np.uint32))
    s = -1 if b >> 31 else 1
    e = (b >> 23) & 0xff
    f = b & 0x7fffff
    if e == 0:
        return s, f, -149
    return s, f | 0x800000, e - 150


def piece_summary(x, sgn, e):
    """Integer summary of one piece for an accumulator of sign `sgn` in binade e.  Returns None (unsafe) or
    [(delta, mn, mx) for parity 0, 1]."""
    out = []
    for parity in (0, 1):
        d, mn, mx = 0, 0, 0
        for v in x:
            if not np.isfinite(v):
                return None
            s, M, ex = decompose(v)
            s *= sgn
            if M == 0:
                continue
            sh = (e - 23) - ex
            if sh <= 0:
                if -sh > 2:
                    return None
                inc = M << (-sh)
            elif sh >= 25:
                inc = 0
            else:
                q, rem, half = M >> sh, M & ((1 << sh) - 1), 1 << (sh - 1)
                if rem > half:
                    q += 1
                elif rem == half:                       # tie: to the even result
                    if (parity + d + s * q) & 1:
                        q += 1
                inc = q
            d += s * inc
            mn, mx = min(mn, d), max(mx, d)
        out.append((d, mn, mx))
    return out


def spec_sum(x, piece=64, stats=None):
    x = np.asarray(x, np.float32)
    n = len(x)
    bounds = list(range(0, n, piece)) + [n]
    est = np.concatenate([[0.0], np.cumsum([x[a:b].astype(np.float64).sum() for a, b in zip(bounds[:-1], bounds[1:])])])
    acc = np.float32(x[0])
    for v in x[1:bounds[1]]:
        acc = np.float32(acc + v)
    replays = 0
    for p in range(1, len(bounds) - 1):
        a, b = bounds[p], bounds[p + 1]
        ok = False
        g = est[p]
        if np.isfinite(g) and g != 0:
            sgn = 1 if g > 0 else -1
            e = int(np.floor(np.log2(abs(g))))
            summ = piece_summary(x[a:b], sgn, e) if -100 <= e <= 100 else None
            bits = int(np.float32(acc).view(np.uint32))
            if summ is not None and np.isfinite(acc) and acc != 0 and (1 if acc > 0 else -1) == sgn and ((bits >> 23) & 0xff) - 127 == e:
                ai = (bits & 0x7fffff) | 0x800000
                d, mn, mx = summ[ai & 1]
                if ai + mn >= (1 << 23) + 1 and ai + mx <= (1 << 24) - 2:
                    ni = ai + d
                    nb = (bits & 0xff800000) | (ni & 0x7fffff)
                    acc = np.uint32(nb).view(np.float32)
                    ok = True
        if not ok:
            replays += 1
            for v in x[a:b]:
                acc = np.float32(acc + v)
    if stats is not None:
        stats.append((len(bounds) - 1, replays))
    return np.float32(acc)


if __name__ == "__main__":
    rs = np.random.RandomState(0)
    cases = []
    for n in (1, 5, 63, 64, 65, 1000, 5000):
        cases.append(rs.standard_normal(n).astype(np.float32))
        cases.append((rs.standard_normal(n) + 0.3).astype(np.float32) * 7)
        cases.append(rs.poisson(3, n).astype(np.float32))
        cases.append(-np.abs(rs.standard_normal(n)).astype(np.float32) * 1e3)
        v = rs.poisson(2, n).astype(np.float32)                                  # deviance-like terms
        with np.errstate(all='ignore'):
            cases.append(np.nan_to_num(v * np.log(v / 2.2)).astype(np.float32))
        # ties: halves of the ulp of a sum near 2^k
        cases.append(np.concatenate([[np.float32(2 ** 20)], (rs.randint(-3, 4, n) * 2.0 ** -4).astype(np.float32)]))
        cases.append(np.concatenate([[np.float32(-2 ** 12 - 1)], (rs.randint(-5, 6, n) * 2.0 ** -12).astype(np.float32)]))
        cases.append(np.concatenate([[np.float32(1.0)], (rs.randint(-1, 2, n) * 2.0 ** -24).astype(np.float32)]))
    bad = 0
    st = []
    for x in cases:
        a, b = seq_sum(x), spec_sum(x, stats=st)
        if a.view(np.uint32) != b.view(np.uint32):
            bad += 1
            print("MISMATCH", len(x), a, b)
    print(f"{len(cases)} cases, {bad} mismatches; pieces/replays of the long cases: {st[-8:]}")
