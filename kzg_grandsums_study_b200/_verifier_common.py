"""Shared host driver of the two verifiers -- drop-in for reference src/grandsum/mset_eq_kzg_verifier.js:9-313 and
src/grandproduct/mset_eq_kzg_verifier.js:9-299.  Host code only (about ten G1 scalar multiplications, one pairing
product); field values are handled as Python integers, group elements through host_bn254."""
import re

from . import host_bn254 as hb
from .Keccak256Transcript import Keccak256Transcript
from .curve import R, getHostCurve
from .ptau_utils import readPTauHost


def verify(kind, pTauFilename, proof, nBits, logger=None):
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

    # ---- STEP 6: [D]_1  (:116-121 / grand-product :103-128)
    if gs:
        D1 = hb.g1_sub(hb.g1_mul(pts["S"], (L1xi - alpha * fxig % R * txig + u) % R), hb.g1_mul(pts["Q"], ZHxi))
    else:
        fxig = (fxi + gamma) % R
        if isSelected:
            fxig = ((fxig - 1) * val["selFxi"] + 1) % R
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
