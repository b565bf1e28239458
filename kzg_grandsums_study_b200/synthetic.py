"""Deterministic synthetic inputs, numpy-vectorised.

PRNG = SplitMix64(seed); one Fr candidate = 4 consecutive outputs as little-endian u64 limbs with the top
two bits of limb 3 cleared; candidates >= r are rejected.  Output layout: standard-form 32 B LE per element
(what the provers take for F / T, like Evaluations.getRandomEvals -> Fr.random(), evaluations.js:51-57).
"""
import numpy as np

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_R_LIMBS = np.array([(R >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64_block(seed, start, count):
    """outputs number start+1 .. start+count of SplitMix64(seed) (1-based like successive next() calls)"""
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _less_than_r(limbs):
    """limbs: (m, 4) uint64 little-endian; elementwise < r"""
    lt = np.zeros(len(limbs), dtype=bool)
    eq = np.ones(len(limbs), dtype=bool)
    for i in (3, 2, 1, 0):
        lt |= eq & (limbs[:, i] < _R_LIMBS[i])
        eq &= limbs[:, i] == _R_LIMBS[i]
    return lt


def random_fr_std(seed, n):
    """n field elements as an (n, 4) uint64 array == n*32 bytes of standard-form LE scalars"""
    out = np.empty((n, 4), dtype=np.uint64)
    filled = 0
    consumed = 0
    while filled < n:
        want = max(1024, int((n - filled) * 1.4) + 16)
        cand = splitmix64_block(seed, consumed, 4 * want).reshape(want, 4).copy()
        cand[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)
        ok = _less_than_r(cand)
        good = cand[ok]
        take = min(len(good), n - filled)
        if take < len(good):
            # stop consuming right after the candidate that produced the last accepted element
            last = np.flatnonzero(ok)[take - 1]
            consumed += 4 * (int(last) + 1)
        else:
            consumed += 4 * want
        out[filled:filled + take] = good[:take]
        filled += take
    return out


def tau_from_seed(seed):
    a = random_fr_std(seed, 1)[0]
    return sum(int(a[i]) << (64 * i) for i in range(4))


def permutation(seed, n):
    """Fisher-Yates driven by SplitMix64(seed): for i = n-1 .. 1: j = next() % (i+1); swap(i, j)"""
    r = splitmix64_block(seed, 0, max(n - 1, 0))
    p = np.arange(n, dtype=np.int64)
    for t, i in enumerate(range(n - 1, 0, -1)):
        j = int(r[t] % np.uint64(i + 1))
        p[i], p[j] = p[j], p[i]
    return p
