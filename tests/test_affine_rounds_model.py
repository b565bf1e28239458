"""CPU model of the batched-affine bucket rounds of csrc/msm.cu (msm_aff_forward / fq_batch_inverse / msm_aff_backward),
with Python integers and the oracle's group law as the judge.

One round adds the entries of every bucket pairwise BY POSITION (entries 2i, 2i+1 -> point i of the next list, an odd
last entry is copied); all denominators of the round are inverted together by Montgomery's trick, with d = 1 standing in
for the pairs that need no quotient (operand at infinity, P + (-P)) and d = 2y for a doubling.  The model states that
contract -- same classification, same formulas -- and the test checks what the CUDA path relies on: the per-bucket sums
are unchanged, a bucket of k entries holds ceil(k / 2^r) points after r rounds, and the result of the MSM
(G1.multiExpAffine, reference src/polynomial/polynomial.js:1112) is the same whatever the number of rounds.
The second half models the linked pieces of a host-scalar MSM (MsmReduceLink / MsmCarry: a later piece opens its buckets
with what the earlier ones left, empty buckets are picked up by the reduction) the same way.
"""
import random

from oracle.py import bn254 as bn

Q = bn.Q
INF = None


def _classify(p1, p2):
    """-> (kind, denominator) exactly as aff_classify in msm.cu"""
    if p1 is INF:
        return "take2", 1
    if p2 is INF:
        return "take1", 1
    if p1[0] == p2[0]:
        if p1[1] == p2[1] and p1[1] != 0:
            return "double", 2 * p1[1] % Q
        return "inf", 1
    return "add", (p2[0] - p1[0]) % Q


def affine_round(buckets):
    """buckets: list of lists of affine points (None = infinity) -> the next lists, with ONE field inversion"""
    pairs = [(b, i) for b, lst in enumerate(buckets) for i in range(len(lst) // 2)]
    kinds, dens = [], []
    for b, i in pairs:
        kind, d = _classify(buckets[b][2 * i], buckets[b][2 * i + 1])
        kinds.append(kind)
        dens.append(d)
    # forward pass: exclusive prefix products; one inversion; backward pass peels the inverses off
    prefix, acc = [], 1
    for d in dens:
        prefix.append(acc)
        acc = acc * d % Q
    s = pow(acc, Q - 2, Q)
    out = [[INF] * ((len(lst) + 1) // 2) for lst in buckets]
    for j in range(len(pairs) - 1, -1, -1):
        b, i = pairs[j]
        p1, p2 = buckets[b][2 * i], buckets[b][2 * i + 1]
        inv = s * prefix[j] % Q
        s = s * dens[j] % Q
        kind = kinds[j]
        if kind in ("add", "double"):
            num = (p2[1] - p1[1]) % Q if kind == "add" else 3 * p1[0] * p1[0] % Q
            lam = num * inv % Q
            x3 = (lam * lam - p1[0] - p2[0]) % Q
            out[b][i] = (x3, (lam * (p1[0] - x3) - p1[1]) % Q)
        elif kind == "take1":
            out[b][i] = p1
        elif kind == "take2":
            out[b][i] = p2
        else:
            out[b][i] = INF
    for b, lst in enumerate(buckets):
        if len(lst) % 2:
            out[b][-1] = lst[-1]
    return out


def _sum(points):
    acc = INF
    for p in points:
        acc = bn.g1_add(acc, p)
    return acc


def test_rounds_keep_the_bucket_sums_and_halve_the_lists():
    rng = random.Random(7)
    G = bn.G1_GEN
    pts = [bn.g1_mul_gen(rng.randrange(1, bn.R)) for _ in range(40)]
    buckets = [
        [rng.choice(pts) for _ in range(k)] for k in (0, 1, 2, 3, 7, 8, 13)
    ] + [
        [G] * 9,                                   # doublings, level after level
        [G, bn.g1_neg(G)] * 4 + [pts[0]],          # cancellations, then infinity meets a point
        [INF, pts[1], INF, INF, pts[2]],           # operands at infinity
        [pts[3], pts[3], bn.g1_neg(pts[3]), pts[4]],
    ]
    want = [_sum(lst) for lst in buckets]
    cur = buckets
    for r in range(1, 6):
        cur = affine_round(cur)
        for b, lst in enumerate(cur):
            assert len(lst) == -(-len(buckets[b]) // (1 << r)), (r, b)
            assert _sum(lst) == want[b], (r, b)
            assert all(p is INF or bn.g1_is_on_curve(p) for p in lst)


def test_msm_result_does_not_depend_on_the_rounds():
    """sum_b (b + 1) S_b over signed-digit buckets of a small MSM == the oracle's MSM, after 0..4 rounds"""
    rng = random.Random(11)
    n, c = 60, 4
    bases = [bn.g1_mul_gen(rng.randrange(1, bn.R)) for _ in range(n)]
    scalars = [rng.randrange(0, bn.R) for _ in range(n)]
    want = bn.g1_msm(bases, scalars)
    half = 1 << (c - 1)
    nwin = -(-257 // c)
    # the window-table flavour: digit w of point i adds +-(2^(c w) P_i) to bucket |digit| - 1 of ONE bucket set
    buckets = [[] for _ in range(half)]
    for i, s in enumerate(scalars):
        carry = 0
        for w in range(nwin):
            raw = ((s >> (c * w)) & ((1 << c) - 1)) + carry
            carry = 1 if raw > half else 0
            mag = (1 << c) - raw if raw > half else raw
            if mag:
                p = bn.g1_mul(bases[i], 1 << (c * w))
                buckets[mag - 1].append(bn.g1_neg(p) if raw > half else p)
    for rounds in range(5):
        cur = buckets
        for _ in range(rounds):
            cur = affine_round(cur)
        acc = INF
        for b, lst in enumerate(cur):
            acc = bn.g1_add(acc, bn.g1_mul(_sum(lst), b + 1))
        assert acc == want, rounds


# ---- linked pieces of a host-scalar MSM (csrc/msm.cu: MsmReduceLink / MsmCarry) ---------------------------------------
# kzg_srs_msm_host cuts the scalars into pieces whose uploads hide behind compute.  Over the window table all pieces share
# ONE bucket geometry; every piece but the last stops after folding its bucket sums -- one point per NON-EMPTY bucket,
# nothing for an empty one -- and a later piece opens every bucket of its walk with the sum the LATEST earlier piece that
# touched the bucket left for it (which already contains the pieces before that one).  Buckets that are empty in the last
# piece are picked up by the reduction the same way.  The model states that contract.
def _digit_buckets(bases, scalars, first, c, nwin):
    """entries of the points first .. first + len(scalars) - 1, window-table flavour: bucket -> list of points"""
    half = 1 << (c - 1)
    buckets = {}
    for k, s in enumerate(scalars):
        carry = 0
        for w in range(nwin):
            raw = ((s >> (c * w)) & ((1 << c) - 1)) + carry
            carry = 1 if raw > half else 0
            mag = (1 << c) - raw if raw > half else raw
            if mag:
                p = bn.g1_mul(bases[first + k], 1 << (c * w))
                buckets.setdefault(mag - 1, []).append(bn.g1_neg(p) if raw > half else p)
    return buckets


def _carry_point(links, b):
    """carry_point() of msm.cu: the latest earlier piece that holds bucket b decides (it contains the older ones)"""
    for link in reversed(links):
        if b in link:
            return link[b]
    return INF


def linked_pieces_msm(bases, scalars, cuts, c, rounds=0):
    nwin = -(-257 // c)
    half = 1 << (c - 1)
    links = []                                     # per finished piece: {bucket: folded sum} for its non-empty buckets
    bounds = [0] + list(cuts) + [len(scalars)]
    last = None
    for k in range(len(bounds) - 1):
        lo, hi = bounds[k], bounds[k + 1]
        mine = _digit_buckets(bases, scalars[lo:hi], lo, c, nwin)
        keys = sorted(mine)
        lists = [mine[b] for b in keys]
        for _ in range(rounds):                    # the rounds see this piece's entries only
            lists = affine_round(lists)
        sums = {}
        for b, lst in zip(keys, lists):            # the walk: opens the bucket with the carry, then adds its own points
            acc = _carry_point(links, b)
            for p in lst:
                acc = bn.g1_add(acc, p)
            sums[b] = acc
        if k + 1 < len(bounds) - 1:
            links.append(sums)
        else:
            last = sums
    acc = INF                                      # the last piece's reduction: own bucket, else what the earlier ones left
    for b in range(half):
        s_b = last[b] if b in last else _carry_point(links, b)
        acc = bn.g1_add(acc, bn.g1_mul(s_b, b + 1))
    return acc


def test_linked_pieces_merge_without_group_operations():
    rng = random.Random(23)
    n, c = 48, 5
    bases = [bn.g1_mul_gen(rng.randrange(1, bn.R)) for _ in range(n)]
    uniform = [rng.randrange(0, bn.R) for _ in range(n)]
    skew = list(uniform)
    for i in range(0, 12):
        skew[i] = 0                                # the whole first piece contributes nothing
    for i in range(30, n):
        skew[i] = rng.randrange(0, 1 << 10)        # short scalars: most buckets empty in the last piece
    same = [7] * n                                 # one bucket holds everything, in every piece
    cancel = [3 if i % 2 == 0 else bn.R - 3 for i in range(n)]
    for scalars in (uniform, skew, same, cancel):
        want = bn.g1_msm(bases, scalars)
        for cuts in ((), (12,), (12, 30), (1, 2), (20, 20)):   # one, two, three pieces; tiny and empty ones
            for rounds in (0, 2):
                assert linked_pieces_msm(bases, scalars, cuts, c, rounds) == want, (cuts, rounds, scalars[0])
