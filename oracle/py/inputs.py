"""Deterministic synthetic inputs (SURVEY.md 8d / BASELINE.md 4) -- ORACLE-side generator.

PRNG = SplitMix64(seed); an Fr element = 4 consecutive u64 as LE limbs, top two bits of limb 3 cleared
(254 bits), rejection-sampled < r; stored as standard-form 32 B LE (what the provers expect for F/T,
mirroring Evaluations.getRandomEvals -> Fr.random(), evaluations.js:51-57).
The product has its own copy of this generator (kzg_grandsums_study_b200/synthetic.py, numpy-vectorised);
tests check the two agree.
"""
from . import bn254 as bn

_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & _M64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def fr(self):
        while True:
            l0, l1, l2, l3 = self.next(), self.next(), self.next(), self.next()
            x = l0 | (l1 << 64) | (l2 << 128) | ((l3 & 0x3FFFFFFFFFFFFFFF) << 192)
            if x < bn.R:
                return x


def tau_from_seed(seed):
    return SplitMix64(seed).fr()


def random_column(seed, n):
    g = SplitMix64(seed)
    return [g.fr() for _ in range(n)]


def rotate_right(v):
    """T = F rotated right by one (test/mset_eq_kzg_grandsum.test.js:28-30)."""
    return v[-1:] + v[:-1]


def permutation(seed, n):
    g = SplitMix64(seed)
    p = list(range(n))
    for i in range(n - 1, 0, -1):
        j = g.next() % (i + 1)
        p[i], p[j] = p[j], p[i]
    return p


def to_std_bytes(v):
    return bn.fr_vec_to_std_bytes(v)
