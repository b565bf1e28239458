"""GPU: every bulk primitive of the hot path through the C ABI against the CPU oracle (bit-exact)."""
import ctypes as C

import numpy as np
import pytest

from oracle.py import bn254 as bn, inputs, poly as opoly

pytestmark = pytest.mark.gpu
R = bn.R


def mont_bytes(v):
    return bn.fr_vec_to_mont_bytes(v)


def from_mont(b):
    return bn.fr_vec_from_mont_bytes(b)


def col(seed, n):
    return inputs.random_column(seed, n)


def test_selftest(curve):
    curve.check(curve.lib.kzg_selftest(curve.ctx, 1 << 14))


def test_montgomery_conversions(curve):
    v = col(1, 1000) + [0, 1, R - 1]
    std = bn.fr_vec_to_std_bytes(v)
    m = curve.Fr.batchToMontgomery(std)
    assert m.tobytes() == mont_bytes(v)
    assert curve.Fr.batchFromMontgomery(m).tobytes() == std


@pytest.mark.parametrize("log_n", list(range(0, 15)))
def test_ntt_small_vs_oracle(curve, log_n):
    n = 1 << log_n
    a = col(100 + log_n, n)
    fwd = curve.Fr.fft(mont_bytes(a))
    assert from_mont(fwd.tobytes()) == opoly.ntt(a)
    inv = curve.Fr.ifft(mont_bytes(a))
    assert from_mont(inv.tobytes()) == opoly.ntt(a, inverse=True)


@pytest.mark.parametrize("log_n", [15, 16, 17, 18, 20, 22, 24])
def test_ntt_large_properties(curve, log_n):
    """sizes the Python oracle cannot finish: spot-check outputs against Horner evaluation (the definition),
    and the inverse round trip"""
    from kzg_grandsums_study_b200 import synthetic
    n = 1 << log_n
    a = synthetic.random_fr_std(200 + log_n, n)   # any 254-bit residues < r are valid Montgomery elements
    buf = curve.to_device(a.tobytes())
    fwd = curve.Fr.fft(buf)
    back = curve.Fr.ifft(fwd)
    assert back.tobytes() == a.tobytes()
    # out[k] = P(w^k): check a few k with the device Horner evaluation and, independently, a host one
    from kzg_grandsums_study_b200.polynomial import Polynomial
    p = Polynomial(buf, curve)
    w = bn.FR_W[log_n]
    ks = [0, 1, 2, n // 2, n - 1, 12345 % n, (n // 3) | 1]
    out = fwd.tobytes()
    for k in ks:
        x = pow(w, k, R)
        got = out[32 * k:32 * k + 32]
        assert p.evaluate(bn.fr_to_mont_bytes(x)) == got, k
    if log_n <= 16:
        coeffs = from_mont(a.tobytes())
        for k in ks[:3]:
            x = pow(w, k, R)
            acc = 0
            for c in reversed(coeffs):
                acc = (acc * x + c) % R
            assert bn.fr_from_mont_bytes(out[32 * k:32 * k + 32]) == acc


@pytest.mark.parametrize("n_coef,ext", [(8, 1), (8, 2), (8, 4), (5, 4), (1024, 4), (3000, 2)])
def test_extend_ntt(curve, n_coef, ext):
    """Evaluations.fromPolynomial(p, extension): zero-pad to nextpow2(len)*ext and NTT (evaluations.js:12-21)"""
    from kzg_grandsums_study_b200.polynomial import Polynomial, Evaluations
    a = col(7, n_coef)
    p = Polynomial(mont_bytes(a), curve)
    ev = Evaluations.fromPolynomial(p, ext, curve)
    size = (1 << (n_coef - 1).bit_length()) * ext
    assert ev.length() == size
    assert from_mont(ev.tobytes()) == opoly.ntt(a + [0] * (size - n_coef))


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 2048, 2049, 5000, 70000])
def test_batch_inverse(curve, n):
    v = col(3, n)
    for i in range(0, n, 7):
        v[i] = 0          # zero maps to zero (ffjavascript batchInverse, grandsum.js:41)
    if n > 40:
        for i in range(32, 40):
            v[i] = 0
    out = curve.Fr.batchInverse(mont_bytes(v))
    assert from_mont(out.tobytes()) == opoly.batch_inverse(v)


@pytest.mark.parametrize("n", [2, 4, 256, 2048, 4096, 1 << 15])
@pytest.mark.parametrize("selected", [False, True])
def test_grand_builders(curve, n, selected):
    """ComputeSGrandSumPolynomial / ComputeZGrandProductPolynomial (grandsum.js:6-62, grandproduct.js:6-57)"""
    from kzg_grandsums_study_b200.grandsum import ComputeSGrandSumPolynomial
    from kzg_grandsums_study_b200.grandproduct import ComputeZGrandProductPolynomial
    from kzg_grandsums_study_b200.polynomial import Evaluations
    from oracle.py import protocol as pr
    f = col(11, n)
    gamma = inputs.SplitMix64(77).fr()
    if selected:
        t = inputs.rotate_right(f)
        sel_f = [1] * (n - 1) + [0]
        sel_t = [0] + [1] * (n - 1)
    else:
        perm = inputs.permutation(9, n)
        t = [f[perm[i]] for i in range(n)]
        sel_f = sel_t = [1] * n
    ev = lambda v: Evaluations(mont_bytes(v), curve)
    sF, sT = (ev(sel_f), ev(sel_t)) if selected else (None, None)
    s = ComputeSGrandSumPolynomial(ev(f), ev(t), sF, sT, bn.fr_to_mont_bytes(gamma), curve)
    assert from_mont(s.tobytes()) == pr.compute_s_grand_sum(f, t, sel_f, sel_t, gamma).coef
    z = ComputeZGrandProductPolynomial(ev(f), ev(t), sF, sT, selected, bn.fr_to_mont_bytes(gamma), curve)
    assert from_mont(z.tobytes()) == pr.compute_z_grand_product(f, t, sel_f, sel_t, gamma).coef


def test_grand_builders_reject_unequal_multisets(curve):
    from kzg_grandsums_study_b200 import KzgError
    from kzg_grandsums_study_b200.grandsum import ComputeSGrandSumPolynomial
    from kzg_grandsums_study_b200.grandproduct import ComputeZGrandProductPolynomial
    from kzg_grandsums_study_b200.polynomial import Evaluations
    n = 64
    f = col(11, n)
    t = list(f)
    t[5] = (t[5] + 1) % R
    ev = lambda v: Evaluations(mont_bytes(v), curve)
    g = bn.fr_to_mont_bytes(12345)
    with pytest.raises(KzgError, match="The grand-sum polynomial S is not well calculated"):
        ComputeSGrandSumPolynomial(ev(f), ev(t), None, None, g, curve)
    with pytest.raises(KzgError, match="The grand-product polynomial Z is not well calculated"):
        ComputeZGrandProductPolynomial(ev(f), ev(t), None, None, False, g, curve)


def test_polynomial_api_vs_oracle(curve):
    """the Polynomial methods the provers use, against the oracle's restatement of polynomial.js"""
    from kzg_grandsums_study_b200.polynomial import Polynomial
    P = lambda v: Polynomial(mont_bytes(v), curve)
    O = opoly.Polynomial
    a, b = col(1, 64), col(2, 128)
    x = inputs.SplitMix64(5).fr()
    xm = bn.fr_to_mont_bytes(x)
    assert from_mont(P(a).add(P(b)).tobytes()) == O(list(a)).add(O(list(b))).coef
    assert from_mont(P(b).add(P(a)).tobytes()) == O(list(b)).add(O(list(a))).coef
    assert from_mont(P(a).sub(P(b)).tobytes()) == O(list(a)).sub(O(list(b))).coef
    assert from_mont(P(a).mulScalar(xm).tobytes()) == O(list(a)).mul_scalar(x).coef
    assert from_mont(P(a).addScalar(xm).tobytes()) == O(list(a)).add_scalar(x).coef
    assert from_mont(P(a).subScalar(xm).tobytes()) == O(list(a)).sub_scalar(x).coef
    assert bn.fr_from_mont_bytes(P(b).evaluate(xm)) == O(list(b)).evaluate(x)
    assert from_mont(P(a).shiftOmega().tobytes()) == O(list(a)).shift_omega().coef
    assert from_mont(P(a).multiply(P(b)).tobytes()) == O(list(a)).multiply(O(list(b))).coef
    assert from_mont(Polynomial.Lagrange1(5, curve).tobytes()) == O.lagrange1(5).coef
    assert P(a + [0, 0, 0]).degree() == 63 and P([0] * 8).degree() == 0 and P([0, 5, 0]).degree() == 1
    # evaluate known answer of test/polynomial.test.js:117-124: 2 + 4x + 2x^2 ... restated: p = [2,4,2] at x=... use oracle
    small = [1, 2, 3, 4]
    assert bn.fr_from_mont_bytes(P(small).evaluate(bn.fr_to_mont_bytes(2))) == 1 + 4 + 12 + 32
    # divByXSubValue: exact division and the "does not divide" error
    from kzg_grandsums_study_b200 import KzgError
    for n in (2, 8, 2048, 5000):
        c = col(3, n)
        val = O(list(c)).evaluate(x)
        c0 = list(c)
        c0[0] = (c0[0] - val) % R
        assert from_mont(P(c0).divByXSubValue(xm).tobytes()) == O(list(c0)).div_by_x_sub_value(x).coef
        with pytest.raises(KzgError, match="Polynomial does not divide"):
            P(c).divByXSubValue(xm)
    # divZh: build (X^n - 1) * q and divide back
    n = 64
    q = col(4, 2 * n - 2) + [0, 0]
    prod = [0] * (4 * n)
    for i, c in enumerate(q):
        prod[i + n] = (prod[i + n] + c) % R
        prod[i] = (prod[i] - c) % R
    got = P(prod).divZh(n)
    ref = O(list(prod)).div_zh(n)
    assert from_mont(got.tobytes()) == ref.coef
    prod[0] = (prod[0] + 1) % R
    with pytest.raises(KzgError, match="Polynomial is not divisible"):
        P(prod).divZh(n)


def test_evaluations_api(curve):
    from kzg_grandsums_study_b200.polynomial import Evaluations
    one = Evaluations.getOneEvals(8, curve)
    zero = Evaluations.getZeroEvals(8, curve)
    assert one.isAllOnes() and not one.isAllZeros() and zero.isAllZeros() and not zero.isAllOnes()
    dev_one = Evaluations(curve.to_device(one.eval), curve)
    assert dev_one.isAllOnes() and not dev_one.isAllZeros()
    dev_one.setEvaluation(3, curve.Fr.zero)
    assert not dev_one.isAllOnes()
    assert dev_one.getEvaluation(3) == curve.Fr.zero and dev_one.getEvaluation(2) == curve.Fr.one
    assert one.length() == 8
    r = Evaluations.getRandomEvals(4, curve)
    assert r.length() == 4 and all(int.from_bytes(r.getEvaluation(i), "little") < R for i in range(4))
    with pytest.raises(IndexError):
        one.getEvaluation(8)


def test_srs_generator_and_ptau_roundtrip(curve, tau, ptau_factory):
    """the synthetic SRS [tau^i]_1 and its .ptau file: prefix against the oracle, header through both readers"""
    from kzg_grandsums_study_b200.ptau_utils import readPTauHeader, readTauG2
    from oracle.py import ptau as opt
    path = ptau_factory(8)
    hdr = readPTauHeader(path)
    assert hdr["power"] == 8 and hdr["ceremonyPower"] == 8
    sections = opt.read_sections(path)
    assert opt.read_ptau_header(path, sections)[0] == 8
    raw = opt.read_tau_g1(path, sections, 512)
    t = 1
    for i in range(40):
        assert raw[64 * i:64 * i + 64] == bn.g1_to_bytes(bn.g1_mul_gen(t)), i
        t = t * tau % R
    i = 511
    assert raw[64 * i:64 * i + 64] == bn.g1_to_bytes(bn.g1_mul_gen(pow(tau, i, R)))
    assert readTauG2(path, curve) == bn.g2_to_bytes(bn.g2_mul(bn.G2_GEN, tau))
    assert opt.read_tau_g2(path, sections) == readTauG2(path, curve)


def test_ptau_header_errors(curve, tmp_path):
    import struct
    from kzg_grandsums_study_b200 import KzgError
    from kzg_grandsums_study_b200.ptau_utils import readPTauHeader
    hdr = struct.pack("<I", 32) + bn.Q.to_bytes(32, "little") + struct.pack("<II", 4, 4)

    def write(name, sections, magic=b"ptau", version=1):
        p = str(tmp_path / name)
        with open(p, "wb") as f:
            f.write(magic + struct.pack("<II", version, len(sections)))
            for sid, payload in sections:
                f.write(struct.pack("<IQ", sid, len(payload)) + payload)
        return p

    with pytest.raises(KzgError, match="File has no  header"):
        readPTauHeader(write("a.ptau", [(2, b"")]))
    with pytest.raises(KzgError, match="more than one header"):
        readPTauHeader(write("b.ptau", [(1, hdr), (1, hdr)]))
    with pytest.raises(KzgError, match="Invalid PTau header size"):
        readPTauHeader(write("c.ptau", [(1, hdr + b"\0\0\0\0")]))
    with pytest.raises(KzgError):
        readPTauHeader(write("d.ptau", [(1, hdr)], magic=b"zkey"))
    assert readPTauHeader(write("e.ptau", [(1, hdr)]))["power"] == 4


@pytest.mark.parametrize("n", [1, 2, 3, 17, 255, 256, 1000, 4096])
def test_msm_vs_oracle(curve, tau, ptau_factory, n):
    """G1.multiExpAffine + toAffine (polynomial.js:1112-1113) against the oracle's Pippenger and the closed form"""
    from oracle.py import ptau as opt
    path = ptau_factory(11)
    sections = opt.read_sections(path)
    bases = opt.read_tau_g1(path, sections, n)
    scalars = col(50 + n, n)
    if n > 3:
        scalars[1] = 0
        scalars[2] = R - 1
        scalars[3] = 1
    jac = curve.G1.multiExpAffine(bases, bn.fr_vec_to_std_bytes(scalars))
    got = curve.G1.toAffine(jac)
    expect = sum(s * pow(tau, i, R) for i, s in enumerate(scalars)) % R
    assert got == bn.g1_to_bytes(bn.g1_mul_gen(expect))
    if n <= 256:
        pts = [bn.g1_from_bytes(bases[64 * i:64 * i + 64]) for i in range(n)]
        assert got == bn.g1_to_bytes(bn.g1_msm(pts, scalars))


def test_msm_edge_cases(curve):
    G = bn.g1_to_bytes((1, 2))
    negG = bn.g1_to_bytes((1, bn.Q - 2))
    inf = bytes(64)
    std = lambda v: bn.fr_vec_to_std_bytes(v)
    msm = lambda b, s: curve.G1.toAffine(curve.G1.multiExpAffine(b, std(s)))
    assert msm(b"", []) == inf                                  # nPoints == 0 -> G1.zero
    assert msm(G, [0]) == inf
    assert msm(G + negG, [5, 5]) == inf                         # P + (-P)
    assert msm(G + G, [3, 4]) == bn.g1_to_bytes(bn.g1_mul_gen(7))   # equal points -> doubling branch
    assert msm(G + inf + G, [1, 9, 1]) == bn.g1_to_bytes(bn.g1_mul_gen(2))   # infinity among the bases
    assert msm(G * 300, [1] * 300) == bn.g1_to_bytes(bn.g1_mul_gen(300))     # one bucket, many equal points
    assert msm(G, [R - 1]) == negG
    # scalars above r are used as plain integers (ffjavascript reads raw bits)
    raw = ((1 << 256) - 1).to_bytes(32, "little")
    got = curve.G1.toAffine(curve.G1.multiExpAffine(G, raw))
    assert got == bn.g1_to_bytes(bn.g1_mul_gen(((1 << 256) - 1) % R))


@pytest.mark.parametrize("window", [2, 5, 8, 11, 13, 16])
def test_msm_window_independence(curve, tau, ptau_factory, window):
    """the commitment is canonical: every window size gives the same bytes"""
    from oracle.py import ptau as opt
    n = 3000
    path = ptau_factory(11)
    bases = opt.read_tau_g1(path, opt.read_sections(path), n)
    scalars = col(8, n)
    expect = sum(s * pow(tau, i, R) for i, s in enumerate(scalars)) % R
    curve.check(curve.lib.kzg_msm_set_window(curve.ctx, window))
    try:
        got = curve.G1.toAffine(curve.G1.multiExpAffine(bases, bn.fr_vec_to_std_bytes(scalars)))
    finally:
        curve.check(curve.lib.kzg_msm_set_window(curve.ctx, 0))
    assert got == bn.g1_to_bytes(bn.g1_mul_gen(expect))


@pytest.mark.parametrize("log_n", [16, 20])
def test_commit_large_closed_form(curve, tau, log_n):
    """commit(p) == p(tau) * G1 at sizes the oracle's MSM cannot reach (independent known answer)"""
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    from kzg_grandsums_study_b200.polynomial import Polynomial
    n = 1 << log_n
    srs = C.c_void_p()
    curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    try:
        coef = synthetic.random_fr_std(31 + log_n, n)            # treated as Montgomery residues
        p = Polynomial(coef.tobytes(), curve)
        got = p.multiExponentiation(srs)
        p_tau = bn.fr_from_mont_bytes(p.evaluate(bn.fr_to_mont_bytes(tau)))
        assert got == bn.g1_to_bytes(bn.g1_mul_gen(p_tau))
        # linearity: commit(a) + commit(b) == commit(a + b)
        coef2 = synthetic.random_fr_std(77, n)
        q = Polynomial(coef2.tobytes(), curve)
        cq = q.multiExponentiation(srs)
        cs = p.add(q).multiExponentiation(srs)
        assert bn.g1_add(bn.g1_from_bytes(got), bn.g1_from_bytes(cq)) == bn.g1_from_bytes(cs)
    finally:
        curve.lib.kzg_srs_free(curve.ctx, srs)


@pytest.mark.parametrize("n,table_c", [(1, 0), (2, 3), (37, 4), (1000, 0), (1000, 9), (5000, 13), (1 << 16, 0), (1 << 16, 17)])
def test_srs_table_msm(curve, tau, n, table_c):
    """MSM over a resident SRS with the precomputed window table (single bucket set, radix-16 reduction hierarchy)
    == the table-less per-window MSM == the closed form; prefixes and offset slices of the SRS included"""
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    srs = C.c_void_p()
    curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    try:
        scal = synthetic.random_fr_std(900 + n, n)
        sc = [sum(int(scal[i, j]) << (64 * j) for j in range(4)) for i in range(min(n, 2000))]
        buf = curve.to_device(scal.tobytes())
        out_plain, out_tab, out_host = bytearray(64), bytearray(64), bytearray(64)
        curve.check(curve.lib.kzg_srs_msm(curve.ctx, srs, 0, buf.handle, n, as_ptr(out_plain)))
        curve.check(curve.lib.kzg_srs_precompute(curve.ctx, srs, table_c))
        c, w = C.c_uint32(), C.c_uint32()
        curve.check(curve.lib.kzg_msm_geometry(curve.ctx, srs, n, 0, C.byref(c), C.byref(w)))
        assert (table_c == 0 or c.value == table_c) and w.value == -(-257 // c.value)
        curve.check(curve.lib.kzg_srs_msm(curve.ctx, srs, 0, buf.handle, n, as_ptr(out_tab)))
        curve.check(curve.lib.kzg_srs_msm_host(curve.ctx, srs, 0, as_ptr(scal), n, as_ptr(out_host)))
        assert bytes(out_tab) == bytes(out_plain) == bytes(out_host)
        if n <= 2000:
            expect = sum(s * pow(tau, i, R) for i, s in enumerate(sc)) % R
            assert bytes(out_tab) == bn.g1_to_bytes(bn.g1_mul_gen(expect))
        if n >= 37:
            first, m = 5, n - 11          # a slice that starts inside the SRS (the multi-GPU shard shape)
            a, b = bytearray(64), bytearray(64)
            curve.check(curve.lib.kzg_srs_msm(curve.ctx, srs, first, buf.handle, m, as_ptr(a)))
            curve.check(curve.lib.kzg_msm_set_window(curve.ctx, 7))       # forces the table-less path
            curve.check(curve.lib.kzg_srs_msm(curve.ctx, srs, first, buf.handle, m, as_ptr(b)))
            curve.check(curve.lib.kzg_msm_set_window(curve.ctx, 0))
            assert bytes(a) == bytes(b)
    finally:
        curve.check(curve.lib.kzg_msm_set_window(curve.ctx, 0))
        curve.lib.kzg_srs_free(curve.ctx, srs)


def test_srs_table_edge_cases(curve):
    """table over an SRS that contains infinity, repeated and opposite points; skewed scalars (one overfull bucket)"""
    from kzg_grandsums_study_b200._lib import as_ptr
    G = bn.g1_to_bytes((1, 2))
    negG = bn.g1_to_bytes((1, bn.Q - 2))
    inf = bytes(64)
    pts = (G + inf + G + negG) * 300
    n = 1200
    srs = C.c_void_p()
    curve.check(curve.lib.kzg_srs_from_host(curve.ctx, as_ptr(pts), n, C.byref(srs)))
    try:
        for table_c in (0, 5):
            curve.check(curve.lib.kzg_srs_precompute(curve.ctx, srs, table_c))
            for scalars in ([7] * n, inputs.random_column(3, n), [0] * n, [R - 1] * n, [1 << 253] * n):
                coeffs = [1, 0, 1, -1] * 300
                expect = sum(s * c for s, c in zip(scalars, coeffs)) % R
                buf = curve.to_device(bn.fr_vec_to_std_bytes(scalars))
                out = bytearray(64)
                curve.check(curve.lib.kzg_srs_msm(curve.ctx, srs, 0, buf.handle, n, as_ptr(out)))
                assert bytes(out) == bn.g1_to_bytes(bn.g1_mul_gen(expect)), (table_c, scalars[0])
    finally:
        curve.lib.kzg_srs_free(curve.ctx, srs)


@pytest.mark.parametrize("table_c", [2, 3, 4, 6, 7, 9, 10, 12, 14, 15, 17])
def test_msm_reduction_geometries(curve, tau, table_c, monkeypatch):
    """the bucket reduction (folded partials -> level 0 -> LO x HI tail in quad-lane arithmetic) for every shape of the
    tail: bucket sets of 2 ... 2^16 buckets, every level-0 radix, sparse and dense buckets, skewed scalars"""
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    n = 1500
    srs = C.c_void_p()
    curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    try:
        curve.check(curve.lib.kzg_srs_precompute(curve.ctx, srs, table_c))
        scal = synthetic.random_fr_std(4200 + table_c, n)
        cases = [[sum(int(scal[i, j]) << (64 * j) for j in range(4)) for i in range(n)],
                 [(i % 3) + 1 for i in range(n)],                       # three buckets hold everything
                 [((1 << table_c) - 1) << (table_c * (i % 5)) for i in range(n)]]  # top digits: carries into the next window
        for k0 in ("0", "1", "2", "3", "5"):
            curve.set_option("red_k0", int(k0))
            for scalars in cases:
                expect = sum(s * pow(tau, i, R) for i, s in enumerate(scalars)) % R
                buf = curve.to_device(bn.fr_vec_to_std_bytes(scalars))
                out = bytearray(64)
                curve.check(curve.lib.kzg_srs_msm(curve.ctx, srs, 0, buf.handle, n, as_ptr(out)))
                assert bytes(out) == bn.g1_to_bytes(bn.g1_mul_gen(expect)), (table_c, k0)
    finally:
        curve.set_option("red_k0", -1)
        curve.lib.kzg_srs_free(curve.ctx, srs)


@pytest.mark.parametrize("flavour,c", [("table", 4), ("table", 9), ("table", 13), ("table", 16), ("table", 20),
                                        ("raw", 5), ("raw", 11), ("raw", 16)])
def test_msm_partition_sort(curve, tau, flavour, c, monkeypatch):
    """the two-level partition sort (shared-memory histograms; the path of every MSM above ~2^16 points), forced on at a
    size the oracle can check: uniform, skewed (all entries in one bucket / one partition) and Montgomery-source scalars"""
    from kzg_grandsums_study_b200 import synthetic
    from kzg_grandsums_study_b200._lib import as_ptr
    n = 2600
    srs = C.c_void_p()
    curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    try:
        if flavour == "table":
            curve.check(curve.lib.kzg_srs_precompute(curve.ctx, srs, c))
        else:
            curve.check(curve.lib.kzg_msm_set_window(curve.ctx, c))
        scal = synthetic.random_fr_std(5100 + c, n)
        cases = [[sum(int(scal[i, j]) << (64 * j) for j in range(4)) for i in range(n)],
                 [5] * n, [0] * n, [R - 1 - (i % 2) for i in range(n)]]
        for part_sort in ("1", "0"):
            curve.set_option("part_sort", int(part_sort))
            for scalars in cases:
                expect = sum(s * pow(tau, i, R) for i, s in enumerate(scalars)) % R
                want = bn.g1_to_bytes(bn.g1_mul_gen(expect))
                buf = curve.to_device(bn.fr_vec_to_std_bytes(scalars))
                out = bytearray(64)
                curve.check(curve.lib.kzg_srs_msm(curve.ctx, srs, 0, buf.handle, n, as_ptr(out)))
                assert bytes(out) == want, (flavour, c, part_sort, scalars[0])
                # Montgomery source (commit of a polynomial): fp_from_mont fused into the digit kernels
                mont = curve.to_device(bn.fr_vec_to_mont_bytes(scalars))
                curve.check(curve.lib.kzg_commit(curve.ctx, srs, mont.handle, as_ptr(out)))
                assert bytes(out) == want, (flavour, c, part_sort, "commit")
    finally:
        curve.set_option("part_sort", -1)
        curve.check(curve.lib.kzg_msm_set_window(curve.ctx, 0))
        curve.lib.kzg_srs_free(curve.ctx, srs)


@pytest.mark.parametrize("part_sort", ["0", "1"])
def test_msm_huge_buckets(curve, tau, part_sort, monkeypatch):
    """buckets with hundreds of partial sums (one scalar value repeated: the block-tier collapse) next to buckets with a
    few dozen (warp tier) and ordinary ones, through both sort schemes"""
    from kzg_grandsums_study_b200._lib import as_ptr
    n = 9000
    srs = C.c_void_p()
    curve.check(curve.lib.kzg_srs_generate(curve.ctx, as_ptr(tau.to_bytes(32, "little")), n, C.byref(srs)))
    try:
        curve.check(curve.lib.kzg_srs_precompute(curve.ctx, srs, 12))
        curve.set_option("part_sort", int(part_sort))
        rnd = inputs.random_column(77, n)
        scalars = [3 if i < 6000 else (5 << 12) if i < 6600 else rnd[i] for i in range(n)]
        expect = sum(s * pow(tau, i, R) for i, s in enumerate(scalars)) % R
        buf = curve.to_device(bn.fr_vec_to_std_bytes(scalars))
        out = bytearray(64)
        curve.check(curve.lib.kzg_srs_msm(curve.ctx, srs, 0, buf.handle, n, as_ptr(out)))
        assert bytes(out) == bn.g1_to_bytes(bn.g1_mul_gen(expect))
    finally:
        curve.set_option("part_sort", -1)
        curve.lib.kzg_srs_free(curve.ctx, srs)


@pytest.mark.parametrize("tile", [4, 8])
def test_ntt_tile_widths(curve, tile):
    """both tile widths of the pass kernels (256-byte rows / 128-byte rows) against the oracle, sizes with an odd and
    an even number of stages per pass, forward, inverse and zero-padded"""
    try:
        curve.set_option("ntt_tile", tile)
        for log_n in (11, 12, 13, 14):
            n = 1 << log_n
            a = col(900 + log_n, n)
            assert from_mont(curve.Fr.fft(mont_bytes(a)).tobytes()) == opoly.ntt(a)
            assert from_mont(curve.Fr.ifft(mont_bytes(a)).tobytes()) == opoly.ntt(a, inverse=True)
        from kzg_grandsums_study_b200.polynomial import Evaluations, Polynomial
        coefs = col(951, 3000)
        ev = Evaluations.fromPolynomial(Polynomial(mont_bytes(coefs), curve), 4, curve)
        assert from_mont(ev.eval.tobytes()) == opoly.ntt(coefs + [0] * (16384 - 3000))
    finally:
        curve.set_option("ntt_tile", -1)


def test_ntt_direct_twiddle_table(curve):
    """transforms of 2^17 .. 2^24 points take the first pass boundary's twiddles from the 512 MB direct table
    (ntt_big_table = 1, default) or from the hi x lo composite (0): same bytes, forward and inverse; 2^17 also against
    the oracle"""
    from kzg_grandsums_study_b200 import synthetic
    try:
        for log_n in (17, 19, 22):
            n = 1 << log_n
            a = synthetic.random_fr_std(400 + log_n, n).tobytes()
            buf = curve.to_device(a)
            got = {}
            for flag in (2, 0):                      # 2: the table from 2^17 points on (default 1: from 2^21)
                curve.set_option("ntt_big_table", flag)
                got[flag] = (curve.Fr.fft(buf).tobytes(), curve.Fr.ifft(buf).tobytes())
            assert got[0] == got[2], log_n
            assert curve.Fr.ifft(curve.Fr.fft(buf)).tobytes() == a
            if log_n == 17:
                vals = from_mont(a)
                assert from_mont(got[2][0]) == opoly.ntt(vals)
    finally:
        curve.set_option("ntt_big_table", -1)
