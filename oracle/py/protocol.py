"""Grand-sum / grand-product multiset-equality KZG provers and verifiers -- ORACLE, test infrastructure only.

CPU restatement (Python big ints) of
  reference src/grandsum/mset_eq_kzg_prover.js:12-435      -> grandsum_prover
  reference src/grandsum/grandsum.js:6-62                   -> compute_s_grand_sum
  reference src/grandsum/mset_eq_kzg_verifier.js:9-313      -> grandsum_verifier
  reference src/grandproduct/mset_eq_kzg_prover.js:12-415   -> grandproduct_prover
  reference src/grandproduct/grandproduct.js:6-57           -> compute_z_grand_product
  reference src/grandproduct/mset_eq_kzg_verifier.js:9-299  -> grandproduct_verifier
  reference src/polynomial/polynomial_utils.js:1-19         -> zh_eval / l1_eval

Byte conventions at the boundary are the reference's: F/T inputs are 32 B standard-form LE per element
(they go through Fr.batchToMontgomery, prover.js:147-148); selectors are 32 B Montgomery-LE (compared
bytewise against Fr.one / Fr.zero, evaluations.js:118-129); the proof holds 64 B Montgomery-LE affine
commitments and 32 B Montgomery-LE evaluations, in the reference's key insertion order.

The pairing check of the verifiers (verifier.js:182) is done either with a real optimal-ate pairing
(oracle/py/pairing.py) or, when tau is known, through the trapdoor identity B - tau*A = O.
"""
from collections import OrderedDict

from . import bn254 as bn
from . import ptau as pt
from .poly import Polynomial, Evaluations, batch_inverse, ntt

R = bn.R


# ------------------------------------------------------------------------------------------------
# polynomial_utils.js
# ------------------------------------------------------------------------------------------------

def zh_eval(x, nbits):                                # polynomial_utils.js:1-10
    xn = x
    for _ in range(nbits):
        xn = xn * xn % R
    return (xn - 1) % R


def l1_eval(x, zhx, nbits):                           # polynomial_utils.js:12-19
    n = (1 << nbits) % R
    return zhx * bn.fr_inv(n * (x - 1) % R) % R


# ------------------------------------------------------------------------------------------------
# grandsum.js / grandproduct.js
# ------------------------------------------------------------------------------------------------

def compute_s_grand_sum(ev_f, ev_t, sel_f, sel_t, gamma):
    n = len(ev_f)
    num = [0] * n
    den = [0] * n
    for i in range(n):                                # grandsum.js:21-38
        f = (ev_f[i] + gamma) % R
        t = (ev_t[i] + gamma) % R
        num[(i + 1) % n] = (t * sel_f[i] - f * sel_t[i]) % R
        den[(i + 1) % n] = f * t % R
    den = batch_inverse(den)                          # grandsum.js:41
    last = 0
    for i in range(n):                                # grandsum.js:44-51
        j = (i + 1) % n
        last = (num[j] * den[j] + last) % R
        num[j] = last
    if num[0] != 0:                                   # grandsum.js:55-57
        raise ValueError("The grand-sum polynomial S is not well calculated")
    return Polynomial.from_evaluations(num)           # grandsum.js:61


def compute_z_grand_product(ev_f, ev_t, sel_f, sel_t, gamma):
    n = len(ev_f)
    num = [1] * n
    den = [1] * n
    for i in range(n):                                # grandproduct.js:21-33
        a = (ev_f[i] + gamma) % R
        b = (ev_t[i] + gamma) % R
        num[(i + 1) % n] = (sel_f[i] * (a - 1) + 1) % R
        den[(i + 1) % n] = (sel_t[i] * (b - 1) + 1) % R
    den = batch_inverse(den)                          # grandproduct.js:36
    last = 1
    for i in range(n):                                # grandproduct.js:39-46
        j = (i + 1) % n
        last = num[j] * den[j] % R * last % R
        num[j] = last
    if num[0] != 1:                                   # grandproduct.js:50-52
        raise ValueError("The grand-product polynomial Z is not well calculated")
    return Polynomial.from_evaluations(num)           # grandproduct.js:56


# ------------------------------------------------------------------------------------------------
# commitments
# ------------------------------------------------------------------------------------------------

class Srs:
    """SRS as read by the prover (prover.js:83-85): list of affine points decoded from section 2."""

    def __init__(self, path, n_points):
        self.sections = pt.read_sections(path)
        self.power, _ = pt.read_ptau_header(path, self.sections)
        raw = pt.read_tau_g1(path, self.sections, n_points)
        self.points = [bn.g1_from_bytes(raw[i:i + 64]) for i in range(0, len(raw) - len(raw) % 64, 64)]

    def commit(self, p):
        """Polynomial.multiExponentiation (polynomial.js:1106-1115) + commit (prover.js:432-434)."""
        n = p.degree() + 1
        return bn.g1_to_bytes(bn.g1_msm(self.points[:n], p.coef[:n]))


class TrapdoorSrs:
    """Closed form commit(p) = p(tau) * G1 -- an independent known answer for any size."""

    def __init__(self, tau, power):
        self.tau = tau % R
        self.power = power

    def commit(self, p):
        return bn.g1_to_bytes(bn.g1_mul_gen(p.evaluate(self.tau)))


# ------------------------------------------------------------------------------------------------
# shared prover front matter (prover.js:22-81)
# ------------------------------------------------------------------------------------------------

def _normalise_inputs(evals_fs, evals_ts, evals_sel_f, evals_sel_t, srs_power):
    if not isinstance(evals_fs, (list, tuple)):
        evals_fs = [evals_fs]
    if not isinstance(evals_ts, (list, tuple)):
        evals_ts = [evals_ts]
    if len(evals_fs) != len(evals_ts):
        raise ValueError("The lengths of the two vector multisets must be the same.")
    n_pols = len(evals_fs)
    if n_pols == 0:
        raise ValueError("The number of multisets must be greater than 0.")
    for i in range(n_pols):
        if len(evals_fs[i]) != len(evals_ts[i]):
            raise ValueError("The %d-th multiset buffers must have the same length." % i)
        if len(evals_fs[i]) != len(evals_fs[0]):
            raise ValueError("The multiset buffers must all have the same length.")
    n = len(evals_fs[0]) // 32
    if evals_sel_f is None:
        evals_sel_f = bn.FR_ONE_BYTES * n
    if evals_sel_t is None:
        evals_sel_t = bn.FR_ONE_BYTES * n
    if len(evals_sel_f) != len(evals_sel_t):
        raise ValueError("The selection buffers must have the same length.")
    if len(evals_sel_f) != len(evals_fs[0]):
        raise ValueError("The selection buffers must have the same length as the multiset buffers.")
    is_selected = not (evals_sel_f == bn.FR_ONE_BYTES * n and evals_sel_t == bn.FR_ONE_BYTES * n)
    nbits = (n - 1).bit_length() if n > 1 else 0
    if n != 1 << nbits:
        raise ValueError("Polynomial length must be a power of two.")
    if srs_power < nbits:
        raise ValueError("The Powers of Tau file is not sufficiently large to commit the polynomials.")
    fs = [bn.fr_vec_from_std_bytes(b) for b in evals_fs]     # batchToMontgomery of standard-form input
    ts = [bn.fr_vec_from_std_bytes(b) for b in evals_ts]
    sel_f = bn.fr_vec_from_mont_bytes(evals_sel_f)            # selectors are already Montgomery
    sel_t = bn.fr_vec_from_mont_bytes(evals_sel_t)
    return fs, ts, sel_f, sel_t, is_selected, n_pols, nbits, n


def _fe(x):
    return bn.fr_to_mont_bytes(x)


def _ch(transcript):
    return bn.fr_from_mont_bytes(transcript.get_challenge())


# ------------------------------------------------------------------------------------------------
# grand-sum prover
# ------------------------------------------------------------------------------------------------

def grandsum_prover(srs, evals_fs, evals_ts, evals_sel_f=None, evals_sel_t=None, trace=None):
    fs, ts, ev_sel_f, ev_sel_t, is_selected, k, nbits, n = _normalise_inputs(
        evals_fs, evals_ts, evals_sel_f, evals_sel_t, srs.power)
    is_vector = k > 1
    proof = {"evaluations": OrderedDict(), "commitments": OrderedDict()}
    C, E = proof["commitments"], proof["evaluations"]
    ch = {}
    tr = pt.Keccak256Transcript()
    w = bn.FR_W[nbits]

    # ---- round 1 (prover.js:144-179)
    pol_fs = [Polynomial.from_evaluations(v) for v in fs]
    pol_ts = [Polynomial.from_evaluations(v) for v in ts]
    for i in range(k):
        C["F%d" % i if is_vector else "F"] = srs.commit(pol_fs[i])
        C["T%d" % i if is_vector else "T"] = srs.commit(pol_ts[i])
    if is_selected:
        sel_f = Polynomial.from_evaluations(ev_sel_f)
        sel_t = Polynomial.from_evaluations(ev_sel_t)
        C["selF"] = srs.commit(sel_f)
        C["selT"] = srs.commit(sel_t)

    # ---- round 2 (prover.js:181-231)
    for i in range(k):
        tr.add_pol_commitment(C["F%d" % i if is_vector else "F"])
        tr.add_pol_commitment(C["T%d" % i if is_vector else "T"])
    if is_selected:
        tr.add_pol_commitment(C["selF"])
        tr.add_pol_commitment(C["selT"])
    if is_vector:
        ch["beta"] = _ch(tr)
        tr.add_field_element(_fe(ch["beta"]))
    ch["gamma"] = _ch(tr)
    if is_vector:
        pol_f = Polynomial.zero(n)
        pol_t = Polynomial.zero(n)
        for i in range(k - 1, -1, -1):
            pol_f.mul_scalar(ch["beta"]).add(pol_fs[i])
            pol_t.mul_scalar(ch["beta"]).add(pol_ts[i])
        ev_f = Evaluations.from_polynomial(pol_f, 1).eval
        ev_t = Evaluations.from_polynomial(pol_t, 1).eval
    else:
        pol_f, pol_t = pol_fs[0], pol_ts[0]
        ev_f, ev_t = fs[0], ts[0]
    pol_s = compute_s_grand_sum(ev_f, ev_t, ev_sel_f, ev_sel_t, ch["gamma"])
    C["S"] = srs.commit(pol_s)

    # ---- round 3 (prover.js:233-286)
    tr.add_field_element(_fe(ch["gamma"]))
    tr.add_pol_commitment(C["S"])
    ch["alpha"] = _ch(tr)
    alpha, gamma = ch["alpha"], ch["gamma"]
    pol_q = Polynomial.zero(n)
    if is_selected:
        b1 = sel_t.clone().multiply(sel_t.clone())
        pol_q.add(sel_t.clone().sub(b1)).mul_scalar(alpha)
        b2 = sel_f.clone().multiply(sel_f.clone())
        pol_q.add(sel_f.clone().sub(b2)).mul_scalar(alpha)
    q1 = pol_s.clone().shift_omega()
    q1.sub(pol_s)
    f_gamma = pol_f.clone().add_scalar(gamma)
    t_gamma = pol_t.clone().add_scalar(gamma)
    q1.multiply(f_gamma)
    q1.multiply(t_gamma)
    if is_selected:
        sfg = sel_f.clone().multiply(t_gamma)
        stg = sel_t.clone().multiply(f_gamma)
        q1.add(stg)
        q1.sub(sfg)
    else:
        q1.add(pol_f)
        q1.sub(pol_t)
    pol_q.add(q1).mul_scalar(alpha)
    q2 = pol_s.clone().multiply(Polynomial.lagrange1(nbits))
    pol_q.add(q2)
    # the reference divides a length-4n buffer; pad so div_zh sees whole extensions
    ext_len = n * max(2, -(-len(pol_q.coef) // n))
    pol_q.coef = pol_q.coef + [0] * (ext_len - len(pol_q.coef))
    pol_q.div_zh(n)
    C["Q"] = srs.commit(pol_q)

    # ---- round 4 (prover.js:288-318)
    tr.add_field_element(_fe(alpha))
    tr.add_pol_commitment(C["Q"])
    ch["xi"] = _ch(tr)
    xi = ch["xi"]
    for i in range(k):
        E["f%dxi" % i if is_vector else "fxi"] = _fe(pol_fs[i].evaluate(xi))
        E["t%dxi" % i if is_vector else "txi"] = _fe(pol_ts[i].evaluate(xi))
    if is_selected:
        E["selFxi"] = _fe(sel_f.evaluate(xi))
        E["selTxi"] = _fe(sel_t.evaluate(xi))
    sxiw = pol_s.evaluate(xi * w % R)
    E["sxiw"] = _fe(sxiw)

    # ---- round 5 (prover.js:320-413)
    tr.add_field_element(_fe(xi))
    for i in range(k):
        tr.add_field_element(E["f%dxi" % i if is_vector else "fxi"])
        tr.add_field_element(E["t%dxi" % i if is_vector else "txi"])
    if is_selected:
        tr.add_field_element(E["selFxi"])
        tr.add_field_element(E["selTxi"])
    tr.add_field_element(E["sxiw"])
    ch["v"] = _ch(tr)
    v = ch["v"]
    zhxi = zh_eval(xi, nbits)
    l1xi = l1_eval(xi, zhxi, nbits)

    pol_r = Polynomial.zero(n)
    if is_selected:
        self_xi = bn.fr_from_mont_bytes(E["selFxi"])
        selt_xi = bn.fr_from_mont_bytes(E["selTxi"])
        pol_r.add_scalar((selt_xi - selt_xi * selt_xi) % R).mul_scalar(alpha)
        pol_r.add_scalar((self_xi - self_xi * self_xi) % R).mul_scalar(alpha)
    r1 = pol_s.clone().mul_scalar(R - 1).add_scalar(sxiw)
    fxi = pol_f.evaluate(xi)
    txi = pol_t.evaluate(xi)
    fxig = (fxi + gamma) % R
    txig = (txi + gamma) % R
    r1.mul_scalar(fxig)
    r1.mul_scalar(txig)
    if is_selected:
        r1.add_scalar(selt_xi * fxig % R)
        r1.sub_scalar(self_xi * txig % R)
    else:
        r1.add_scalar(fxi)
        r1.sub_scalar(txi)
    pol_r.add(r1).mul_scalar(alpha)
    pol_r.add(pol_s.clone().mul_scalar(l1xi))
    pol_r.sub(pol_q.clone().mul_scalar(zhxi))

    wxi = Polynomial.zero(n)
    if is_selected:
        wxi.add(sel_t.clone().sub_scalar(selt_xi))
        wxi.mul_scalar(v).add(sel_f.clone().sub_scalar(self_xi))
    for i in range(k - 1, -1, -1):
        e = bn.fr_from_mont_bytes(E["t%dxi" % i if is_vector else "txi"])
        wxi.mul_scalar(v).add(pol_ts[i].clone().sub_scalar(e))
    for i in range(k - 1, -1, -1):
        e = bn.fr_from_mont_bytes(E["f%dxi" % i if is_vector else "fxi"])
        wxi.mul_scalar(v).add(pol_fs[i].clone().sub_scalar(e))
    wxi.mul_scalar(v).add(pol_r.clone())
    wxi.div_by_x_sub_value(xi)
    wxiw = pol_s.clone().sub_scalar(sxiw)
    wxiw.div_by_x_sub_value(xi * w % R)
    C["Wxi"] = srs.commit(wxi)
    C["Wxiw"] = srs.commit(wxiw)
    if trace is not None:
        trace.update(challenges=ch, pol_s=pol_s, pol_q=pol_q, wxi=wxi, wxiw=wxiw,
                     pol_fs=pol_fs, pol_ts=pol_ts)
    return proof


# ------------------------------------------------------------------------------------------------
# grand-product prover
# ------------------------------------------------------------------------------------------------

def grandproduct_prover(srs, evals_fs, evals_ts, evals_sel_f=None, evals_sel_t=None, trace=None):
    fs, ts, ev_sel_f, ev_sel_t, is_selected, k, nbits, n = _normalise_inputs(
        evals_fs, evals_ts, evals_sel_f, evals_sel_t, srs.power)
    is_vector = k > 1
    proof = {"evaluations": OrderedDict(), "commitments": OrderedDict()}
    C, E = proof["commitments"], proof["evaluations"]
    ch = {}
    tr = pt.Keccak256Transcript()
    w = bn.FR_W[nbits]

    # ---- round 1 (grandproduct/prover.js:144-179)
    pol_fs = [Polynomial.from_evaluations(v) for v in fs]
    pol_ts = [Polynomial.from_evaluations(v) for v in ts]
    for i in range(k):
        C["F%d" % i if is_vector else "F"] = srs.commit(pol_fs[i])
        C["T%d" % i if is_vector else "T"] = srs.commit(pol_ts[i])
    if is_selected:
        sel_f = Polynomial.from_evaluations(ev_sel_f)
        sel_t = Polynomial.from_evaluations(ev_sel_t)
        C["selF"] = srs.commit(sel_f)
        C["selT"] = srs.commit(sel_t)

    # ---- round 2 (:181-231)
    for i in range(k):
        tr.add_pol_commitment(C["F%d" % i if is_vector else "F"])
        tr.add_pol_commitment(C["T%d" % i if is_vector else "T"])
    if is_selected:
        tr.add_pol_commitment(C["selF"])
        tr.add_pol_commitment(C["selT"])
    if is_vector:
        ch["beta"] = _ch(tr)
        tr.add_field_element(_fe(ch["beta"]))
    ch["gamma"] = _ch(tr)
    if is_vector:
        pol_f = Polynomial.zero(n)
        pol_t = Polynomial.zero(n)
        for i in range(k - 1, -1, -1):
            pol_f.mul_scalar(ch["beta"]).add(pol_fs[i])
            pol_t.mul_scalar(ch["beta"]).add(pol_ts[i])
        ev_f = Evaluations.from_polynomial(pol_f, 1).eval
        ev_t = Evaluations.from_polynomial(pol_t, 1).eval
    else:
        pol_f, pol_t = pol_fs[0].clone(), pol_ts[0].clone()
        ev_f, ev_t = fs[0], ts[0]
    pol_z = compute_z_grand_product(ev_f, ev_t, ev_sel_f, ev_sel_t, ch["gamma"])
    C["Z"] = srs.commit(pol_z)

    # ---- round 3 (:233-287)
    tr.add_field_element(_fe(ch["gamma"]))
    tr.add_pol_commitment(C["Z"])
    ch["alpha"] = _ch(tr)
    alpha, gamma = ch["alpha"], ch["gamma"]
    pol_q = Polynomial.zero(n)
    if is_selected:
        b1 = sel_t.clone().multiply(sel_t.clone())
        pol_q.add(sel_t.clone().sub(b1)).mul_scalar(alpha)
        b2 = sel_f.clone().multiply(sel_f.clone())
        pol_q.add(sel_f.clone().sub(b2)).mul_scalar(alpha)
    q1 = pol_z.clone().shift_omega()
    q2 = pol_z.clone()
    f_gamma = pol_f.clone().add_scalar(gamma)
    t_gamma = pol_t.clone().add_scalar(gamma)
    if is_selected:
        t_gamma.sub_scalar(1).multiply(sel_t.clone()).add_scalar(1)
        q1.multiply(t_gamma)
        f_gamma.sub_scalar(1).multiply(sel_f.clone()).add_scalar(1)
        q2.multiply(f_gamma)
    else:
        q1.multiply(t_gamma)
        q2.multiply(f_gamma)
    q1.sub(q2)
    pol_q.add(q1).mul_scalar(alpha)
    q3 = pol_z.clone().sub_scalar(1).multiply(Polynomial.lagrange1(nbits))
    pol_q.add(q3)
    ext_len = n * max(2, -(-len(pol_q.coef) // n))
    pol_q.coef = pol_q.coef + [0] * (ext_len - len(pol_q.coef))
    pol_q.div_zh(n)
    C["Q"] = srs.commit(pol_q)

    # ---- round 4 (:289-315)
    tr.add_field_element(_fe(alpha))
    tr.add_pol_commitment(C["Q"])
    ch["xi"] = _ch(tr)
    xi = ch["xi"]
    for i in range(k):
        E["f%dxi" % i if is_vector else "fxi"] = _fe(pol_fs[i].evaluate(xi))
    if is_selected:
        E["selFxi"] = _fe(sel_f.evaluate(xi))
        E["selTxi"] = _fe(sel_t.evaluate(xi))
    zxiw = pol_z.evaluate(xi * w % R)
    E["zxiw"] = _fe(zxiw)

    # ---- round 5 (:317-410)
    tr.add_field_element(_fe(xi))
    for i in range(k):
        tr.add_field_element(E["f%dxi" % i if is_vector else "fxi"])
    if is_selected:
        tr.add_field_element(E["selFxi"])
        tr.add_field_element(E["selTxi"])
    tr.add_field_element(E["zxiw"])
    ch["v"] = _ch(tr)
    v = ch["v"]
    zhxi = zh_eval(xi, nbits)
    l1xi = l1_eval(xi, zhxi, nbits)

    pol_r = Polynomial.zero(n)
    if is_selected:
        self_xi = bn.fr_from_mont_bytes(E["selFxi"])
        selt_xi = bn.fr_from_mont_bytes(E["selTxi"])
        pol_r.add_scalar((selt_xi - selt_xi * selt_xi) % R).mul_scalar(alpha)
        pol_r.add_scalar((self_xi - self_xi * self_xi) % R).mul_scalar(alpha)
    r1 = Polynomial.zero(n)
    fxi = pol_f.evaluate(xi)
    fxig = (fxi + gamma) % R
    txig = pol_t.add_scalar(gamma)                    # mutates polT in place (:354)
    if is_selected:
        fxig = (fxig - 1) % R
        txig = txig.sub_scalar(1)
        self_g = (self_xi * fxig + 1) % R
        selt_g = txig.mul_scalar(selt_xi).add_scalar(1)
        selt_g.mul_scalar(zxiw)
        r1.add(selt_g)
        r1.sub(pol_z.clone().mul_scalar(self_g))
    else:
        txig.mul_scalar(zxiw)
        r1.add(txig)
        r1.sub(pol_z.clone().mul_scalar(fxig))
    pol_r.add(r1).mul_scalar(alpha)
    pol_r.add(pol_z.clone().sub_scalar(1).mul_scalar(l1xi))
    pol_r.sub(pol_q.clone().mul_scalar(zhxi))

    wxi = Polynomial.zero(n)
    if is_selected:
        wxi.add(sel_t.clone().sub_scalar(selt_xi))
        wxi.mul_scalar(v).add(sel_f.clone().sub_scalar(self_xi))
    for i in range(k - 1, -1, -1):
        e = bn.fr_from_mont_bytes(E["f%dxi" % i if is_vector else "fxi"])
        wxi.mul_scalar(v).add(pol_fs[i].clone().sub_scalar(e))
    wxi.mul_scalar(v).add(pol_r.clone())
    wxi.div_by_x_sub_value(xi)
    wxiw = pol_z.clone().sub_scalar(zxiw)
    wxiw.div_by_x_sub_value(xi * w % R)
    C["Wxi"] = srs.commit(wxi)
    C["Wxiw"] = srs.commit(wxiw)
    if trace is not None:
        trace.update(challenges=ch, pol_z=pol_z, pol_q=pol_q, wxi=wxi, wxiw=wxiw)
    return proof


# ------------------------------------------------------------------------------------------------
# verifiers
# ------------------------------------------------------------------------------------------------

def _verifier_challenges(proof, k, is_vector, is_selected, acc_name, with_t):
    """verifier.js:246-312 (grand-sum) / grandproduct verifier.js:236-298."""
    C, E = proof["commitments"], proof["evaluations"]
    tr = pt.Keccak256Transcript()
    ch = {}
    for i in range(k):
        tr.add_pol_commitment(C["F%d" % i if is_vector else "F"])
        tr.add_pol_commitment(C["T%d" % i if is_vector else "T"])
    if is_selected:
        tr.add_pol_commitment(C["selF"])
        tr.add_pol_commitment(C["selT"])
    if is_vector:
        ch["beta"] = _ch(tr)
        tr.add_field_element(_fe(ch["beta"]))
    ch["gamma"] = _ch(tr)
    tr.add_field_element(_fe(ch["gamma"]))
    tr.add_pol_commitment(C[acc_name])
    ch["alpha"] = _ch(tr)
    tr.add_field_element(_fe(ch["alpha"]))
    tr.add_pol_commitment(C["Q"])
    ch["xi"] = _ch(tr)
    tr.add_field_element(_fe(ch["xi"]))
    for i in range(k):
        tr.add_field_element(E["f%dxi" % i if is_vector else "fxi"])
        if with_t:
            tr.add_field_element(E["t%dxi" % i if is_vector else "txi"])
    if is_selected:
        tr.add_field_element(E["selFxi"])
        tr.add_field_element(E["selTxi"])
    tr.add_field_element(E["sxiw" if acc_name == "S" else "zxiw"])
    ch["v"] = _ch(tr)
    tr.add_field_element(_fe(ch["v"]))
    tr.add_pol_commitment(C["Wxi"])
    tr.add_pol_commitment(C["Wxiw"])
    ch["u"] = _ch(tr)
    return ch


def _pt(b):
    return bn.g1_from_bytes(b)


def _g1_lin(terms):
    """sum_i k_i * P_i (O(1) many terms)."""
    acc = None
    for kk, P in terms:
        acc = bn.g1_add(acc, bn.g1_mul(P, kk)) if acc is not None else bn.g1_mul(P, kk)
    return acc


def _valid_proof_values(proof):
    for c in proof["commitments"].values():           # verifier.js:199-228 (G1.isValid)
        if len(c) != 64:
            return False
        if c != bytes(64):
            if int.from_bytes(c[:32], "little") >= bn.Q or int.from_bytes(c[32:], "little") >= bn.Q:
                return False
            if not bn.g1_is_on_curve(_pt(c)):
                return False
    for e in proof["evaluations"].values():           # verifier.js:230-244
        if int.from_bytes(e, "little") >= R:
            return False
    return True


def _final_check(A, B, tau=None, tau_g2=None):
    """verifier.js:182: e(-A,[tau]_2) * e(B,[1]_2) == 1.  With the trapdoor: B - tau*A == O."""
    if tau is not None:
        return bn.g1_add(B, bn.g1_neg(bn.g1_mul(A, tau))) is None
    from .pairing import pairing_eq
    return pairing_eq(bn.g1_neg(A), tau_g2, B, bn.G2_GEN)


def grandsum_verifier(proof, nbits, tau=None, tau_g2=None, out_challenges=None):
    C, E = proof["commitments"], proof["evaluations"]
    n_fi = len([key for key in C if key[0] == "F" and key[1:2].isdigit()])
    k = n_fi if n_fi > 0 else 1
    is_vector = k > 1
    is_selected = len([key for key in C if key.startswith("selF")]) == 1
    if not _valid_proof_values(proof):
        return False
    ch = _verifier_challenges(proof, k, is_vector, is_selected, "S", True)
    if out_challenges is not None:
        out_challenges.update(ch)
    ev = {key: bn.fr_from_mont_bytes(val) for key, val in E.items()}
    alpha, gamma, xi, v, u = ch["alpha"], ch["gamma"], ch["xi"], ch["v"], ch["u"]
    zhxi = zh_eval(xi, nbits)
    l1xi = l1_eval(xi, zhxi, nbits)
    r0 = 0                                            # verifier.js:78-111
    if is_selected:
        r0 = (r0 + ev["selTxi"] - ev["selTxi"] ** 2) * alpha % R
        r0 = (r0 + ev["selFxi"] - ev["selFxi"] ** 2) * alpha % R
    fxi = txi = 0
    for i in range(k - 1, -1, -1):
        fxi = (fxi * ch.get("beta", 0) + ev["f%dxi" % i if is_vector else "fxi"]) % R
        txi = (txi * ch.get("beta", 0) + ev["t%dxi" % i if is_vector else "txi"]) % R
    fxig, txig = (fxi + gamma) % R, (txi + gamma) % R
    r01 = ev["sxiw"] * fxig % R * txig % R
    if is_selected:
        r01 = (r01 + ev["selTxi"] * fxig - ev["selFxi"] * txig) % R
    else:
        r01 = (r01 + fxi - txi) % R
    r0 = (r0 + r01) * alpha % R
    d1 = (l1xi - alpha * fxig % R * txig + u) % R    # verifier.js:116-121
    D = _g1_lin([(d1, _pt(C["S"])), ((-zhxi) % R, _pt(C["Q"]))])
    F1 = None                                         # verifier.js:126-142
    if is_selected:
        F1 = _pt(C["selT"])
        F1 = bn.g1_add(bn.g1_mul(F1, v), _pt(C["selF"]))
    for i in range(k - 1, -1, -1):
        F1 = bn.g1_add(bn.g1_mul(F1, v), _pt(C["T%d" % i if is_vector else "T"]))
    for i in range(k - 1, -1, -1):
        F1 = bn.g1_add(bn.g1_mul(F1, v), _pt(C["F%d" % i if is_vector else "F"]))
    F1 = bn.g1_add(bn.g1_mul(F1, v), D)
    e1 = 0                                            # verifier.js:147-167
    if is_selected:
        e1 = ev["selTxi"]
        e1 = (e1 * v + ev["selFxi"]) % R
    for i in range(k - 1, -1, -1):
        e1 = (e1 * v + ev["t%dxi" % i if is_vector else "txi"]) % R
    for i in range(k - 1, -1, -1):
        e1 = (e1 * v + ev["f%dxi" % i if is_vector else "fxi"]) % R
    e1 = (e1 * v + u * ev["sxiw"] - r0) % R
    E1 = bn.g1_mul_gen(e1)
    A = bn.g1_add(_pt(C["Wxi"]), bn.g1_mul(_pt(C["Wxiw"]), u))          # verifier.js:172-180
    B = bn.g1_add(_pt(C["Wxi"]), bn.g1_mul(_pt(C["Wxiw"]), u * bn.FR_W[nbits] % R))
    B = bn.g1_mul(B, xi)
    B = bn.g1_add(B, F1)
    B = bn.g1_add(B, bn.g1_neg(E1))
    return _final_check(A, B, tau, tau_g2)


def grandproduct_verifier(proof, nbits, tau=None, tau_g2=None, out_challenges=None):
    C, E = proof["commitments"], proof["evaluations"]
    n_fi = len([key for key in C if key[0] == "F" and key[1:2].isdigit()])
    k = n_fi if n_fi > 0 else 1
    is_vector = k > 1
    is_selected = len([key for key in C if key.startswith("selF")]) == 1
    if not _valid_proof_values(proof):
        return False
    ch = _verifier_challenges(proof, k, is_vector, is_selected, "Z", False)
    if out_challenges is not None:
        out_challenges.update(ch)
    ev = {key: bn.fr_from_mont_bytes(val) for key, val in E.items()}
    alpha, gamma, xi, v, u = ch["alpha"], ch["gamma"], ch["xi"], ch["v"], ch["u"]
    beta = ch.get("beta", 0)
    zhxi = zh_eval(xi, nbits)
    l1xi = l1_eval(xi, zhxi, nbits)
    r0 = 0                                            # grandproduct verifier.js:78-98
    if is_selected:
        r0 = (r0 + ev["selTxi"] - ev["selTxi"] ** 2) * alpha % R
        r0 = (r0 + ev["selFxi"] - ev["selFxi"] ** 2) * alpha % R
    r01 = ev["zxiw"]
    if is_selected:
        r01 = r01 * (((gamma - 1) * ev["selTxi"] + 1) % R) % R
    else:
        r01 = r01 * gamma % R
    r0 = (r0 + r01) * alpha % R
    r0 = (r0 - l1xi) % R
    fxi = 0                                           # :103-128
    for i in range(k - 1, -1, -1):
        fxi = (fxi * beta + ev["f%dxi" % i if is_vector else "fxi"]) % R
    fxig = (fxi + gamma) % R
    if is_selected:
        fxig = ((fxig - 1) * ev["selFxi"] + 1) % R
    d1 = (l1xi - alpha * fxig + u) % R
    D1 = bn.g1_mul(_pt(C["Z"]), d1)
    D2 = None
    for i in range(k - 1, -1, -1):
        D2 = bn.g1_add(bn.g1_mul(D2, beta), _pt(C["T%d" % i if is_vector else "T"]))
    if is_selected:
        D2 = bn.g1_mul(D2, ev["selTxi"])
    D2 = bn.g1_mul(D2, ev["zxiw"])
    D2 = bn.g1_mul(D2, alpha)
    D3 = bn.g1_mul(_pt(C["Q"]), zhxi)
    D = bn.g1_add(bn.g1_add(D1, D2), bn.g1_neg(D3))
    F1 = None                                         # :130-143
    if is_selected:
        F1 = _pt(C["selT"])
        F1 = bn.g1_add(bn.g1_mul(F1, v), _pt(C["selF"]))
    for i in range(k - 1, -1, -1):
        F1 = bn.g1_add(bn.g1_mul(F1, v), _pt(C["F%d" % i if is_vector else "F"]))
    F1 = bn.g1_add(bn.g1_mul(F1, v), D)
    e1 = 0                                            # :145-163
    if is_selected:
        e1 = ev["selTxi"]
        e1 = (e1 * v + ev["selFxi"]) % R
    for i in range(k - 1, -1, -1):
        e1 = (e1 * v + ev["f%dxi" % i if is_vector else "fxi"]) % R
    e1 = (e1 * v + u * ev["zxiw"] - r0) % R
    E1 = bn.g1_mul_gen(e1)
    A = bn.g1_add(_pt(C["Wxi"]), bn.g1_mul(_pt(C["Wxiw"]), u))          # :167-175
    B = bn.g1_mul(_pt(C["Wxiw"]), u * xi % R * bn.FR_W[nbits] % R)
    B = bn.g1_add(bn.g1_mul(_pt(C["Wxi"]), xi), B)
    B = bn.g1_add(B, F1)
    B = bn.g1_add(B, bn.g1_neg(E1))
    return _final_check(A, B, tau, tau_g2)


def proof_bytes(proof):
    """Canonical concatenation for parity: commitments then evaluations, in the
    reference's key-insertion order, raw 64 B / 32 B Montgomery-LE values."""
    return b"".join(proof["commitments"].values()) + b"".join(proof["evaluations"].values())
