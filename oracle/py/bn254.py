"""BN254 (alt_bn128) field / group arithmetic and encodings -- ORACLE, test infrastructure only.

This file is part of the CPU oracle for the KZG grand-sum / grand-product hot path.  It is a
from-scratch restatement with Python big integers; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import it.  The product path (libkzgb200.so) never does.

The arithmetic of the reference lives in the un-vendored npm dependency ffjavascript@0.2.59 /
wasmcurves@0.2.1 (reference package-lock.json:300-309, 905-912).  What is restated here is its
*published behaviour* as pinned down in SURVEY.md Appendix B/C:

  * Fr / Fq elements in memory: 32-byte little-endian Montgomery residues (R = 2^256), fully reduced.
  * G1 affine: x || y (64 B, Montgomery-LE), infinity = 64 zero bytes.  Jacobian: x || y || z (96 B).
  * G1.toRprUncompressed: x || y, 32-byte big-endian standard form (Keccak256Transcript.js:42).
  * Fr.toRprBE: 32-byte big-endian standard form (Keccak256Transcript.js:45).
  * Roots of unity: w[28] = 5^((r-1)/2^28), w[i] = w[i+1]^2.

PARITY PIN: the reference's own tests hold no golden vectors; this oracle is pinned
against public constants (EIP-196 2*G1, keccak256("")), SURVEY.md Appendix B constants and the
Appendix F end-to-end vector.  "Parity unpinned" at the ffjavascript boundary -- see DESIGN.md.
"""

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MONT = 1 << 256
MONT_INV_Q = pow(MONT, -1, Q)
MONT_INV_R = pow(MONT, -1, R)

# ---------------------------------------------------------------------------------------------
# encodings
# ---------------------------------------------------------------------------------------------

def fr_to_mont_bytes(x):
    """standard-form integer -> 32 B Montgomery little-endian (what sits in .coef / .eval buffers)."""
    return ((x % R) * MONT % R).to_bytes(32, "little")


def fr_from_mont_bytes(b):
    return int.from_bytes(b, "little") * MONT_INV_R % R


def fq_to_mont_bytes(x):
    return ((x % Q) * MONT % Q).to_bytes(32, "little")


def fq_from_mont_bytes(b):
    return int.from_bytes(b, "little") * MONT_INV_Q % Q


def fr_std_bytes(x):
    """standard-form 32 B LE (what Fr.random() returns and the provers take as F/T inputs)."""
    return (x % R).to_bytes(32, "little")


def fr_vec_to_mont_bytes(v):
    return b"".join(fr_to_mont_bytes(x) for x in v)


def fr_vec_from_mont_bytes(b):
    return [fr_from_mont_bytes(b[i:i + 32]) for i in range(0, len(b), 32)]


def fr_vec_to_std_bytes(v):
    return b"".join(fr_std_bytes(x) for x in v)


def fr_vec_from_std_bytes(b):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


FR_ONE_BYTES = fr_to_mont_bytes(1)
FR_ZERO_BYTES = bytes(32)

# ---------------------------------------------------------------------------------------------
# roots of unity (ffjavascript: s = 28, non-residue 5)
# ---------------------------------------------------------------------------------------------

FR_S = 28
_w = [0] * (FR_S + 1)
_w[FR_S] = pow(5, (R - 1) >> FR_S, R)
for _i in range(FR_S - 1, -1, -1):
    _w[_i] = _w[_i + 1] * _w[_i + 1] % R
FR_W = _w  # FR_W[k] is a primitive 2^k-th root of unity


def fr_inv(x):
    return pow(x, -1, R) if x % R else 0


# ---------------------------------------------------------------------------------------------
# G1: y^2 = x^3 + 3 over Fq.  Points are None (infinity) or (x, y) affine ints; Jacobian (X, Y, Z).
# ---------------------------------------------------------------------------------------------

G1_GEN = (1, 2)


def g1_is_on_curve(P):
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - 3) % Q == 0


def jac_double(P):
    X, Y, Z = P
    if Z == 0 or Y == 0:
        return (1, 1, 0)
    A = X * X % Q
    B = Y * Y % Q
    C = B * B % Q
    D = 2 * ((X + B) * (X + B) - A - C) % Q
    E = 3 * A % Q
    F = E * E % Q
    X3 = (F - 2 * D) % Q
    Y3 = (E * (D - X3) - 8 * C) % Q
    Z3 = 2 * Y * Z % Q
    return (X3, Y3, Z3)


def jac_add(P, S):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = S
    if Z1 == 0:
        return S
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % Q
    Z2Z2 = Z2 * Z2 % Q
    U1 = X1 * Z2Z2 % Q
    U2 = X2 * Z1Z1 % Q
    S1 = Y1 * Z2 * Z2Z2 % Q
    S2 = Y2 * Z1 * Z1Z1 % Q
    H = (U2 - U1) % Q
    Rr = (S2 - S1) % Q
    if H == 0:
        if Rr == 0:
            return jac_double(P)
        return (1, 1, 0)
    HH = H * H % Q
    HHH = H * HH % Q
    V = U1 * HH % Q
    X3 = (Rr * Rr - HHH - 2 * V) % Q
    Y3 = (Rr * (V - X3) - S1 * HHH) % Q
    Z3 = Z1 * Z2 * H % Q
    return (X3, Y3, Z3)


def jac_add_affine(P, A):
    """Jacobian + affine (mixed add); A is (x, y) or None."""
    if A is None:
        return P
    X1, Y1, Z1 = P
    x2, y2 = A
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % Q
    U2 = x2 * Z1Z1 % Q
    S2 = y2 * Z1 * Z1Z1 % Q
    H = (U2 - X1) % Q
    Rr = (S2 - Y1) % Q
    if H == 0:
        if Rr == 0:
            return jac_double(P)
        return (1, 1, 0)
    HH = H * H % Q
    HHH = H * HH % Q
    V = X1 * HH % Q
    X3 = (Rr * Rr - HHH - 2 * V) % Q
    Y3 = (Rr * (V - X3) - Y1 * HHH) % Q
    Z3 = Z1 * H % Q
    return (X3, Y3, Z3)


def jac_to_affine(P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, Q)
    zi2 = zi * zi % Q
    return (X * zi2 % Q, Y * zi2 * zi % Q)


def g1_neg(P):
    if P is None:
        return None
    return (P[0], (-P[1]) % Q)


def g1_add(P, S):
    a = (P[0], P[1], 1) if P is not None else (1, 1, 0)
    return jac_to_affine(jac_add_affine(a, S))


def g1_mul(P, k):
    """scalar multiplication, k a standard-form integer (G1.timesFr semantics after decoding)."""
    k %= R
    acc = (1, 1, 0)
    if P is None or k == 0:
        return None
    for bit in bin(k)[2:]:
        acc = jac_double(acc)
        if bit == "1":
            acc = jac_add_affine(acc, P)
    return jac_to_affine(acc)


_GEN_TABLE = None


def g1_mul_gen(k):
    """k * G1 with a lazily built fixed-base 8-bit window table (only used to write synthetic SRS)."""
    global _GEN_TABLE
    if _GEN_TABLE is None:
        tab = []
        base = (G1_GEN[0], G1_GEN[1], 1)
        for _ in range(32):
            row = [None]
            acc = (1, 1, 0)
            for _d in range(255):
                acc = jac_add(acc, base)
                row.append(acc)
            # normalise the row to affine so later additions are mixed adds
            tab.append([None] + [jac_to_affine(p) for p in row[1:]])
            for _d in range(8):
                base = jac_double(base)
        _GEN_TABLE = tab
    k %= R
    acc = (1, 1, 0)
    for w in range(32):
        d = (k >> (8 * w)) & 0xFF
        if d:
            acc = jac_add_affine(acc, _GEN_TABLE[w][d])
    return jac_to_affine(acc)


def g1_to_bytes(P):
    """affine point -> 64 B Montgomery-LE (infinity = zeros); the layout of ptau section 2 and of
    proof.commitments[*] (reference polynomial.js:1112-1113: multiExpAffine + toAffine)."""
    if P is None:
        return bytes(64)
    return fq_to_mont_bytes(P[0]) + fq_to_mont_bytes(P[1])


def g1_from_bytes(b):
    if b == bytes(64):
        return None
    return (fq_from_mont_bytes(b[:32]), fq_from_mont_bytes(b[32:64]))


def g1_to_rpr_uncompressed(b):
    """64 B Montgomery-LE affine -> 64 B big-endian standard form (Keccak256Transcript.js:42).
    Infinity: zeros with 0x40 OR-ed into byte 0 [dep, from memory -- SURVEY.md B.3]."""
    P = g1_from_bytes(b)
    if P is None:
        out = bytearray(64)
        out[0] |= 0x40
        return bytes(out)
    return P[0].to_bytes(32, "big") + P[1].to_bytes(32, "big")


def fr_to_rpr_be(b):
    """32 B Montgomery-LE -> 32 B big-endian standard form (Keccak256Transcript.js:45)."""
    return fr_from_mont_bytes(b).to_bytes(32, "big")


# ---------------------------------------------------------------------------------------------
# naive / Pippenger MSM over affine bases (G1.multiExpAffine; result is canonical so the window
# schedule is irrelevant to the bytes).  Scalars are standard-form integers.
# ---------------------------------------------------------------------------------------------

def g1_msm(bases, scalars, c=None):
    n = len(scalars)
    assert len(bases) >= n
    if n == 0:
        return None
    if c is None:
        c = 1 if n < 4 else max(1, min(16, n.bit_length() - 3))
    nwin = (254 + c - 1) // c
    acc = (1, 1, 0)
    mask = (1 << c) - 1
    for w in range(nwin - 1, -1, -1):
        for _ in range(c):
            acc = jac_double(acc)
        buckets = {}
        sh = w * c
        for i in range(n):
            d = (scalars[i] >> sh) & mask
            if d and bases[i] is not None:
                b = buckets.get(d)
                buckets[d] = jac_add_affine(b, bases[i]) if b is not None else (bases[i][0], bases[i][1], 1)
        if buckets:
            run = (1, 1, 0)
            tot = (1, 1, 0)
            top = max(buckets)
            for d in range(top, 0, -1):
                b = buckets.get(d)
                if b is not None:
                    run = jac_add(run, b)
                tot = jac_add(tot, run)
            acc = jac_add(acc, tot)
    return jac_to_affine(acc)


# ---------------------------------------------------------------------------------------------
# G2 (only to write [tau]_2 into a synthetic ptau and for the pairing check of the verifier).
# Fq2 = Fq[u]/(u^2+1); twist y^2 = x^3 + 3/(9+u).
# ---------------------------------------------------------------------------------------------

def fq2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def fq2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def fq2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def fq2_sqr(a):
    return fq2_mul(a, a)


def fq2_scalar(a, k):
    return (a[0] * k % Q, a[1] * k % Q)


def fq2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)


G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
G2_B = fq2_mul((3, 0), fq2_inv((9, 1)))


def g2_is_on_curve(P):
    if P is None:
        return True
    x, y = P
    return fq2_sub(fq2_sqr(y), fq2_add(fq2_mul(fq2_sqr(x), x), G2_B)) == (0, 0)


def g2_add(P, S):
    if P is None:
        return S
    if S is None:
        return P
    (x1, y1), (x2, y2) = P, S
    if x1 == x2:
        if fq2_add(y1, y2) == (0, 0):
            return None
        lam = fq2_mul(fq2_scalar(fq2_sqr(x1), 3), fq2_inv(fq2_scalar(y1, 2)))
    else:
        lam = fq2_mul(fq2_sub(y2, y1), fq2_inv(fq2_sub(x2, x1)))
    x3 = fq2_sub(fq2_sub(fq2_sqr(lam), x1), x2)
    y3 = fq2_sub(fq2_mul(lam, fq2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_mul(P, k):
    k %= R
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, P)
    return acc


def g2_to_bytes(P):
    """128 B = x.c0 || x.c1 || y.c0 || y.c1, Montgomery-LE; infinity = zeros."""
    if P is None:
        return bytes(128)
    (x, y) = P
    return fq_to_mont_bytes(x[0]) + fq_to_mont_bytes(x[1]) + fq_to_mont_bytes(y[0]) + fq_to_mont_bytes(y[1])


def g2_from_bytes(b):
    if b == bytes(128):
        return None
    v = [fq_from_mont_bytes(b[i:i + 32]) for i in range(0, 128, 32)]
    return ((v[0], v[1]), (v[2], v[3]))
