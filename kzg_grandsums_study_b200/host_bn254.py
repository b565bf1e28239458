"""Host-side BN254 group arithmetic and optimal-ate pairing for the VERIFIERS.

The reference's verifiers (src/grandsum/mset_eq_kzg_verifier.js, src/grandproduct/mset_eq_kzg_verifier.js) run on
the host: about ten `G1.timesFr`, a few additions and one `curve.pairingEq` per proof, no loop over n -- out of
scope for the GPU.  This module gives the drop-in verifiers the same host
primitives ffjavascript gives the JavaScript ones (G1.timesFr / add / sub / neg / isValid, curve.pairingEq), with
plain Python integers.  Nothing here is on the prover hot path.

Pairing: optimal ate over Fq12 = Fq[w] / (w^12 - 18 w^6 + 82), G2 untwisted into Fq12, Miller loop with affine line
functions, one final exponentiation for the whole product (pairingEq only needs `product == 1`).
"""
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_MONT = 1 << 256
_MINV_Q = pow(_MONT, -1, Q)
ATE_LOOP_COUNT = 29793968203157093288
LOG_ATE_LOOP_COUNT = 63

G1_GEN = (1, 2)
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


# ---------------------------------------------------------------------------------------------------------
# encodings (ffjavascript buffers: Montgomery little-endian; infinity = zeros)
# ---------------------------------------------------------------------------------------------------------
def fq_from_mont(b):
    return int.from_bytes(bytes(b), "little") * _MINV_Q % Q


def fq_to_mont(x):
    return (x % Q * _MONT % Q).to_bytes(32, "little")


def g1_from_bytes(b):
    b = bytes(b)
    if b == bytes(64):
        return None
    return (fq_from_mont(b[:32]), fq_from_mont(b[32:]))


def g1_bytes_canonical(b):
    """both coordinates are reduced residues (raw little-endian integer < q).  A coordinate + q decodes to the same
    point but hashes differently in the transcript (G1.toRprUncompressed works on the raw bytes), so the verifiers
    refuse such encodings before the Montgomery conversion makes the difference invisible."""
    b = bytes(b)
    return len(b) == 64 and int.from_bytes(b[:32], "little") < Q and int.from_bytes(b[32:], "little") < Q


def g1_to_bytes(P):
    return bytes(64) if P is None else fq_to_mont(P[0]) + fq_to_mont(P[1])


def g2_from_bytes(b):
    """128 B = x.c0 | x.c1 | y.c0 | y.c1, Montgomery-LE (ptau section 3, verifier.js:18-19)"""
    b = bytes(b)
    if b == bytes(128):
        return None
    v = [fq_from_mont(b[32 * i:32 * i + 32]) for i in range(4)]
    return ((v[0], v[1]), (v[2], v[3]))


# ---------------------------------------------------------------------------------------------------------
# G1 (y^2 = x^3 + 3 over Fq), affine with None = infinity
# ---------------------------------------------------------------------------------------------------------
def g1_is_valid(P):
    """G1.isValid: on the curve (the group has cofactor 1), or infinity"""
    if P is None:
        return True
    x, y = P
    return 0 <= x < Q and 0 <= y < Q and (y * y - x * x * x - 3) % Q == 0


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % Q)


def g1_add(P, S):
    if P is None:
        return S
    if S is None:
        return P
    x1, y1 = P
    x2, y2 = S
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        m = 3 * x1 * x1 * pow(2 * y1, -1, Q) % Q
    else:
        m = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (m * m - x1 - x2) % Q
    return (x3, (m * (x1 - x3) - y1) % Q)


def g1_sub(P, S):
    return g1_add(P, g1_neg(S))


def g1_mul(P, k):
    """G1.timesFr(P, k): k is a plain integer here (the callers convert from Montgomery)"""
    k %= R
    acc = None
    add = P
    while k:
        if k & 1:
            acc = g1_add(acc, add)
        add = g1_add(add, add)
        k >>= 1
    return acc


# ---------------------------------------------------------------------------------------------------------
# Fq12 = Fq[w] / (w^12 - 18 w^6 + 82): coefficient lists of length 12
# ---------------------------------------------------------------------------------------------------------
_F12_ONE = [1] + [0] * 11
_F12_ZERO = [0] * 12


def f12_mul(a, b):
    t = [0] * 23
    for i in range(12):
        ai = a[i]
        if ai:
            for j in range(12):
                t[i + j] += ai * b[j]
    for i in range(22, 11, -1):      # w^12 = 18 w^6 - 82
        c = t[i]
        if c:
            t[i - 6] += 18 * c
            t[i - 12] -= 82 * c
    return [x % Q for x in t[:12]]


def f12_add(a, b):
    return [(x + y) % Q for x, y in zip(a, b)]


def f12_sub(a, b):
    return [(x - y) % Q for x, y in zip(a, b)]


def f12_scalar(a, k):
    return [x * k % Q for x in a]


def _deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def _poly_div(a, b):
    """quotient of a by b over Fq (lists, low degree first)"""
    da, db = _deg(a), _deg(b)
    tmp = list(a)
    out = [0] * len(a)
    inv_lead = pow(b[db], -1, Q)
    for i in range(da - db, -1, -1):
        q = tmp[db + i] * inv_lead % Q
        out[i] = q
        if q:
            for c in range(db + 1):
                tmp[c + i] = (tmp[c + i] - q * b[c]) % Q
    return out[:max(da - db + 1, 1)]


def f12_inv(a):
    """extended Euclid in Fq[w] against the modulus"""
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [82, 0, 0, 0, 0, 0, -18 % Q, 0, 0, 0, 0, 0, 1]
    while _deg(low):
        r = _poly_div(high, low)
        r += [0] * (13 - len(r))
        nm, new = list(hm), list(high)
        for i in range(13):
            li, lo = lm[i], low[i]
            if li or lo:
                for j in range(13 - i):
                    if r[j]:
                        nm[i + j] -= li * r[j]
                        new[i + j] -= lo * r[j]
        nm = [x % Q for x in nm]
        new = [x % Q for x in new]
        lm, low, hm, high = nm, new, lm, low
    inv0 = pow(low[0], -1, Q)
    return [x * inv0 % Q for x in lm[:12]]


def f12_pow(a, e):
    out = list(_F12_ONE)
    base = list(a)
    while e:
        if e & 1:
            out = f12_mul(out, base)
        base = f12_mul(base, base)
        e >>= 1
    return out


# ---------------------------------------------------------------------------------------------------------
# points with Fq12 coordinates (the untwisted G2 point and the embedded G1 point), affine, None = infinity
# ---------------------------------------------------------------------------------------------------------
def _twist(P2):
    """E'(Fq2) -> E(Fq12): with Fq2 = Fq[i]/(i^2+1) and i = w^6 - 9, (x, y) -> (x w^2, y w^3)"""
    (x0, x1), (y0, y1) = P2
    nx = [0] * 12
    ny = [0] * 12
    nx[0], nx[6] = (x0 - 9 * x1) % Q, x1
    ny[0], ny[6] = (y0 - 9 * y1) % Q, y1
    w2 = [0, 0, 1] + [0] * 9
    w3 = [0, 0, 0, 1] + [0] * 8
    return (f12_mul(nx, w2), f12_mul(ny, w3))


def _embed_g1(P):
    return ([P[0]] + [0] * 11, [P[1]] + [0] * 11)


def _p12_double(P):
    x, y = P
    m = f12_mul(f12_scalar(f12_mul(x, x), 3), f12_inv(f12_scalar(y, 2)))
    nx = f12_sub(f12_mul(m, m), f12_scalar(x, 2))
    ny = f12_sub(f12_mul(m, f12_sub(x, nx)), y)
    return (nx, ny)


def _p12_add(P, S):
    if P is None:
        return S
    if S is None:
        return P
    x1, y1 = P
    x2, y2 = S
    if x1 == x2:
        return _p12_double(P) if y1 == y2 else None
    m = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
    nx = f12_sub(f12_sub(f12_mul(m, m), x1), x2)
    ny = f12_sub(f12_mul(m, f12_sub(x1, nx)), y1)
    return (nx, ny)


def _linefunc(P1, P2, T):
    """the line through P1 and P2 (tangent if equal) evaluated at T"""
    x1, y1 = P1
    x2, y2 = P2
    xt, yt = T
    if x1 != x2:
        m = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
    elif y1 == y2:
        m = f12_mul(f12_scalar(f12_mul(x1, x1), 3), f12_inv(f12_scalar(y1, 2)))
    else:
        return f12_sub(xt, x1)
    return f12_sub(f12_mul(m, f12_sub(xt, x1)), f12_sub(yt, y1))


def miller_loop(Q2, P1):
    """f_{6u+2, Q}(P) and the two Frobenius lines -- WITHOUT the final exponentiation.  Q2 in G2 (Fq2 affine), P1 in G1."""
    if Q2 is None or P1 is None:
        return list(_F12_ONE)
    Qt = _twist(Q2)
    P = _embed_g1(P1)
    Rp = Qt
    f = list(_F12_ONE)
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = f12_mul(f12_mul(f, f), _linefunc(Rp, Rp, P))
        Rp = _p12_double(Rp)
        if ATE_LOOP_COUNT & (1 << i):
            f = f12_mul(f, _linefunc(Rp, Qt, P))
            Rp = _p12_add(Rp, Qt)
    Q1 = (f12_pow(Qt[0], Q), f12_pow(Qt[1], Q))
    nQ2 = (f12_pow(Q1[0], Q), f12_sub(_F12_ZERO, f12_pow(Q1[1], Q)))
    f = f12_mul(f, _linefunc(Rp, Q1, P))
    Rp = _p12_add(Rp, Q1)
    f = f12_mul(f, _linefunc(Rp, nQ2, P))
    return f


def final_exponentiate(f):
    return f12_pow(f, (Q ** 12 - 1) // R)


def pairing(Q2, P1):
    return final_exponentiate(miller_loop(Q2, P1))


def pairing_eq(A1, A2, B1, B2):
    """curve.pairingEq(A1, A2, B1, B2): e(A1, A2) * e(B1, B2) == 1"""
    f = f12_mul(miller_loop(A2, A1), miller_loop(B2, B1))
    return final_exponentiate(f) == _F12_ONE
