"""Shared host driver of the two verifiers -- drop-in for reference src/grandsum/mset_eq_kzg_verifier.js:9-313 and
src/grandproduct/mset_eq_kzg_verifier.js:9-299.  Host code only (about ten G1 scalar multiplications, one pairing
product); field values are handled as Python integers, group elements through host_bn254."""
import re

from . import host_bn254 as hb
from .Keccak256Transcript import Keccak256Transcript
from .curve import R, getHostCurve
from .ptau_utils import readPTauHost


def _prepare(kind, pTauFilename, proof, nBits, logger=None):
    """steps 1-5 of the reference verifier: validity checks, the Fiat-Shamir challenges, Z_H(xi), L_1(xi), r_0 -- field
    arithmetic only.  Returns the local state (a dict) for the group part, or False when the proof is malformed."""
    gs = kind == "gs"
    hdr = readPTauHost(pTauFilename)                              # verifier.js:12-20
    curve = getHostCurve()
    Fr = curve.Fr
    X2 = hb.g2_from_bytes(hdr["X2"])
    Cm, Ev = proof["commitments"], proof["evaluations"]
    acc = "S" if gs else "Z"
    acc_eval = "sxiw" if gs else "zxiw"

    nFi = len([k for k in Cm if re.match(r"^F\d", k)])            # :23-28
    nPols = nFi if nFi > 0 else 1
    isVector = nPols > 1
    isSelected = len([k for k in Cm if re.match(r"^selF", k)]) == 1
    fname = lambda i: "F%d" % i if isVector else "F"
    tname = lambda i: "T%d" % i if isVector else "T"
    fev = lambda i: "f%dxi" % i if isVector else "fxi"
    tev = lambda i: "t%dxi" % i if isVector else "txi"

    def err(msg):
        if logger:
            logger.error(msg)
        return False

    # ---- STEP 1: commitments are valid G1 elements (:50, validateCommitments)
    names = []
    for i in range(nPols):
        names += [fname(i), tname(i)]
    if isSelected:
        names += ["selF", "selT"]
    names += [acc, "Q", "Wxi", "Wxiw"]
    pts = {}
    for nm in names:
        if nm not in Cm or len(Cm[nm]) != 64:
            return err("missing commitment %s" % nm)
        if not hb.g1_bytes_canonical(Cm[nm]):
            return err("%s is not a canonically encoded G1 element" % nm)
        P = hb.g1_from_bytes(Cm[nm])
        if not hb.g1_is_valid(P):
            return err("%s is not a valid G1 element" % nm)
        pts[nm] = P

    # ---- STEP 2: evaluations are valid field elements (:61, validateEvaluations; the grand-sum verifier does not
    # range-check selFxi / selTxi either, verifier.js:232-244)
    ev_names = []
    for i in range(nPols):
        ev_names.append(fev(i))
        if gs:
            ev_names.append(tev(i))
    ev_names.append(acc_eval)
    for nm in ev_names:
        if nm not in Ev or int.from_bytes(bytes(Ev[nm]), "little") >= R:
            return err("%s is not a valid field element" % nm)
    if isSelected:
        # (not range-checked by the reference, but they must be there: a selected proof without them is malformed, not
        # a crash)
        for nm in ("selFxi", "selTxi"):
            if nm not in Ev or len(bytes(Ev[nm])) != 32:
                return err("missing evaluation %s" % nm)
    val = {k: Fr.toObject(v) for k, v in Ev.items()}

    # ---- STEP 3: challenges (:246-312)
    tr = Keccak256Transcript(curve)
    ch = {}
    for i in range(nPols):
        tr.addPolCommitment(Cm[fname(i)])
        tr.addPolCommitment(Cm[tname(i)])
    if isSelected:
        tr.addPolCommitment(Cm["selF"])
        tr.addPolCommitment(Cm["selT"])
    beta_b = None
    if isVector:
        beta_b = tr.getChallenge()
        tr.addFieldElement(beta_b)
    gamma_b = tr.getChallenge()
    tr.addFieldElement(gamma_b)
    tr.addPolCommitment(Cm[acc])
    alpha_b = tr.getChallenge()
    tr.addFieldElement(alpha_b)
    tr.addPolCommitment(Cm["Q"])
    xi_b = tr.getChallenge()
    tr.addFieldElement(xi_b)
    for i in range(nPols):
        tr.addFieldElement(Ev[fev(i)])
        if gs:
            tr.addFieldElement(Ev[tev(i)])
    if isSelected:
        tr.addFieldElement(Ev["selFxi"])
        tr.addFieldElement(Ev["selTxi"])
    tr.addFieldElement(Ev[acc_eval])
    v_b = tr.getChallenge()
    tr.addFieldElement(v_b)
    tr.addPolCommitment(Cm["Wxi"])
    tr.addPolCommitment(Cm["Wxiw"])
    u_b = tr.getChallenge()
    beta = Fr.toObject(beta_b) if isVector else 0
    gamma, alpha, xi, v, u = (Fr.toObject(b) for b in (gamma_b, alpha_b, xi_b, v_b, u_b))
    ch.update(beta=beta, gamma=gamma, alpha=alpha, xi=xi, v=v, u=u)

    # ---- STEP 4: ZH(xi), L1(xi)  (polynomial_utils.js)
    xn = xi
    for _ in range(nBits):
        xn = xn * xn % R
    ZHxi = (xn - 1) % R
    L1xi = ZHxi * pow((1 << nBits) * (xi - 1) % R, -1, R) % R

    # ---- STEP 5: r0  (:78-111 / grand-product :78-97)
    r0 = 0
    if isSelected:
        st, sf = val["selTxi"], val["selFxi"]
        r0 = (r0 + st - st * st) * alpha % R
        r0 = (r0 + sf - sf * sf) * alpha % R
    fxi = 0
    txi = 0
    for i in range(nPols - 1, -1, -1):
        fxi = (fxi * beta + val[fev(i)]) % R
        if gs:
            txi = (txi * beta + val[tev(i)]) % R
    if gs:
        fxig, txig = (fxi + gamma) % R, (txi + gamma) % R
        r01 = val["sxiw"] * fxig % R * txig % R
        if isSelected:
            r01 = (r01 + val["selTxi"] * fxig - val["selFxi"] * txig) % R
        else:
            r01 = (r01 + fxi - txi) % R
        r0 = (r0 + r01) * alpha % R
    else:
        r01 = val["zxiw"]
        if isSelected:
            r01 = r01 * (((gamma - 1) * val["selTxi"] + 1) % R) % R
        else:
            r01 = r01 * gamma % R
        r0 = ((r0 + r01) * alpha - L1xi) % R

    if gs:
        fxig, txig = (fxi + gamma) % R, (txi + gamma) % R
    else:
        fxig = (fxi + gamma) % R
        if isSelected:
            fxig = ((fxig - 1) * val["selFxi"] + 1) % R
    return locals()


def verify(kind, pTauFilename, proof, nBits, logger=None):
    st = _prepare(kind, pTauFilename, proof, nBits, logger)
    if st is False:
        return False
    gs, isSelected, nPols, pts, val, Fr, X2 = (st[k] for k in ("gs", "isSelected", "nPols", "pts", "val", "Fr", "X2"))
    fname, tname, fev, tev, acc_eval = (st[k] for k in ("fname", "tname", "fev", "tev", "acc_eval"))
    beta, gamma, alpha, xi, v, u = (st["ch"][k] for k in ("beta", "gamma", "alpha", "xi", "v", "u"))
    ZHxi, L1xi, r0, fxi, fxig = (st[k] for k in ("ZHxi", "L1xi", "r0", "fxi", "fxig"))
    txig = st.get("txig")

    # ---- STEP 6: [D]_1  (:116-121 / grand-product :103-128)
    if gs:
        D1 = hb.g1_sub(hb.g1_mul(pts["S"], (L1xi - alpha * fxig % R * txig + u) % R), hb.g1_mul(pts["Q"], ZHxi))
    else:
        D1_1 = hb.g1_mul(pts["Z"], (L1xi - alpha * fxig + u) % R)
        D1_2 = None
        for i in range(nPols - 1, -1, -1):
            D1_2 = hb.g1_add(hb.g1_mul(D1_2, beta), pts[tname(i)])
        if isSelected:
            D1_2 = hb.g1_mul(D1_2, val["selTxi"])
        D1_2 = hb.g1_mul(hb.g1_mul(D1_2, val["zxiw"]), alpha)
        D1 = hb.g1_sub(hb.g1_add(D1_1, D1_2), hb.g1_mul(pts["Q"], ZHxi))

    # ---- STEP 7: [F]_1  (:126-142)
    F1 = None
    if isSelected:
        F1 = hb.g1_add(F1, pts["selT"])
        F1 = hb.g1_add(hb.g1_mul(F1, v), pts["selF"])
    if gs:
        for i in range(nPols - 1, -1, -1):
            F1 = hb.g1_add(hb.g1_mul(F1, v), pts[tname(i)])
    for i in range(nPols - 1, -1, -1):
        F1 = hb.g1_add(hb.g1_mul(F1, v), pts[fname(i)])
    F1 = hb.g1_add(hb.g1_mul(F1, v), D1)

    # ---- STEP 8: [E]_1  (:147-167)
    E = 0
    if isSelected:
        E = (E + val["selTxi"]) % R
        E = (E * v + val["selFxi"]) % R
    if gs:
        for i in range(nPols - 1, -1, -1):
            E = (E * v + val[tev(i)]) % R
    for i in range(nPols - 1, -1, -1):
        E = (E * v + val[fev(i)]) % R
    E = (E * v + u * val[acc_eval]) % R
    E = (E - r0) % R
    E1 = hb.g1_mul(hb.G1_GEN, E)

    # ---- STEP 9: pairing check (:172-182)
    w = Fr.toObject(Fr.w[nBits])
    A = hb.g1_add(pts["Wxi"], hb.g1_mul(pts["Wxiw"], u))
    B = hb.g1_add(pts["Wxi"], hb.g1_mul(pts["Wxiw"], u * w % R))
    B = hb.g1_mul(B, xi)
    B = hb.g1_sub(hb.g1_add(B, F1), E1)
    isValid = hb.pairing_eq(hb.g1_neg(A), X2, B, hb.G2_GEN)
    if logger:
        (logger.info if isValid else logger.error)("> VERIFICATION OK" if isValid else "> VERIFICATION FAILED")
    return isValid



# ---------------------------------------------------------------------------------------------------------------------
# Batched verification.  One proof costs the reference ~10 G1.timesFr and a 2-pairing check
# (verifier.js:116-182); for m proofs that is 10 m scalar multiplications and m pairing products on the host.  Both sides of
# every check are LINEAR in the proof's commitments:
#     A_p = [Wxi] + u [Wxiw]                                  B_p = xi [Wxi] + u xi w [Wxiw] + [F]_p - [E]_p
# so with random weights rho_p the m checks collapse into ONE:  e(-sum rho_p A_p, [tau]_2) e(sum rho_p B_p, [1]_2) = 1,
# i.e. two multi-scalar multiplications over all the commitments of all the proofs (2 m and ~(2k + 6) m points) -- done
# on the device through G1.multiExpAffine (kzg_g1_msm_affine) -- and ONE pairing product on the host.  All-or-nothing:
# it accepts iff every single proof is valid (up to probability m / r); to find the bad one, fall back to verify().
# ---------------------------------------------------------------------------------------------------------------------
def linear_terms(st):
    """the pairing inputs of one proof as linear forms over its commitments: ({name: scalar} for A, {name: scalar} for B,
    scalar of the generator in B).  Mirrors steps 6-9 of verify() term by term."""
    gs, isSelected, nPols, val = st["gs"], st["isSelected"], st["nPols"], st["val"]
    fname, tname, acc, acc_eval, fev, tev = st["fname"], st["tname"], st["acc"], st["acc_eval"], st["fev"], st["tev"]
    beta, gamma, alpha, xi, v, u = (st["ch"][k] for k in ("beta", "gamma", "alpha", "xi", "v", "u"))
    ZHxi, L1xi, r0, fxig = st["ZHxi"], st["L1xi"], st["r0"], st["fxig"]
    B = {}

    def add(name, c):
        B[name] = (B.get(name, 0) + c) % R

    # [D]_1 (step 6), entering [F]_1 with v^0
    if gs:
        add(acc, L1xi - alpha * fxig % R * st["txig"] + u)
    else:
        add(acc, L1xi - alpha * fxig + u)
        c = alpha * val["zxiw"] % R
        if isSelected:
            c = c * val["selTxi"] % R
        bp = 1
        for i in range(nPols):
            add(tname(i), c * bp)
            bp = bp * beta % R
    add("Q", -ZHxi)
    # [F]_1 (step 7): Horner in v -- the element added LAST carries v^1
    vp = v
    for i in range(nPols):
        add(fname(i), vp)
        vp = vp * v % R
    if gs:
        for i in range(nPols):
            add(tname(i), vp)
            vp = vp * v % R
    if isSelected:
        add("selF", vp)
        vp = vp * v % R
        add("selT", vp)
    # [E]_1 (step 8)
    E = 0
    if isSelected:
        E = (E + val["selTxi"]) % R
        E = (E * v + val["selFxi"]) % R
    if gs:
        for i in range(nPols - 1, -1, -1):
            E = (E * v + val[tev(i)]) % R
    for i in range(nPols - 1, -1, -1):
        E = (E * v + val[fev(i)]) % R
    E = (E * v + u * val[acc_eval]) % R
    E = (E - r0) % R
    # step 9
    w = st["Fr"].toObject(st["Fr"].w[st["nBits"]])
    add("Wxi", xi)
    add("Wxiw", u * xi % R * w)
    A = {"Wxi": 1, "Wxiw": u}
    return A, B, (-E) % R


def verify_batch(kind, pTauFilename, proofs, nBits, curve=None, msm=None, logger=None):
    """True iff EVERY proof of the list verifies (same pTau, same nBits).  `msm(bases_bytes, scalars_std_bytes) -> 64 B
    affine` defaults to the device MSM of `curve` (getCurveFromName: a GPU is required); tests inject a host MSM."""
    if not proofs:
        return True
    if msm is None:
        if curve is None:
            from .curve import getCurveFromName
            curve = getCurveFromName("bn128")
        msm = lambda bases, scalars: curve.G1.toAffine(curve.G1.multiExpAffine(bases, scalars))
    states = []
    for proof in proofs:
        st = _prepare(kind, pTauFilename, proof, nBits, logger)
        if st is False:
            return False
        states.append(st)
    # weights: rho_1 = 1, rho_p = Keccak over every proof's last challenge and a counter (Fiat-Shamir over the batch)
    host = getHostCurve()
    tr = Keccak256Transcript(host)
    for st in states:
        tr.addFieldElement(host.Fr.e(st["ch"]["u"]))
    basesA, scalA, basesB, scalB = [], [], [], []
    gen_scalar = 0
    rho = 1
    for idx, st in enumerate(states):
        if idx > 0:
            tr.addFieldElement(host.Fr.e(idx))
            rho = host.Fr.toObject(tr.getChallenge())
        A, B, e = linear_terms(st)
        Cm = st["Cm"]
        for name, c in A.items():
            basesA.append(bytes(Cm[name]))
            scalA.append(c * rho % R)
        for name, c in B.items():
            basesB.append(bytes(Cm[name]))
            scalB.append(c * rho % R)
        gen_scalar = (gen_scalar + e * rho) % R
    basesB.append(hb.g1_to_bytes(hb.G1_GEN))
    scalB.append(gen_scalar)
    to_std = lambda xs: b"".join(int(x).to_bytes(32, "little") for x in xs)
    A_tot = hb.g1_from_bytes(msm(b"".join(basesA), to_std(scalA)))
    B_tot = hb.g1_from_bytes(msm(b"".join(basesB), to_std(scalB)))
    ok = hb.pairing_eq(hb.g1_neg(A_tot), states[0]["X2"], B_tot, hb.G2_GEN)
    if logger:
        (logger.info if ok else logger.error)("> BATCH VERIFICATION OK (%d proofs)" % len(proofs) if ok else "> BATCH VERIFICATION FAILED")
    return ok
