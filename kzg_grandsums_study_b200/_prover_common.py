"""Shared host driver of the two provers: argument normalisation and checks of the reference
(src/grandsum/mset_eq_kzg_prover.js:12-81 == src/grandproduct/mset_eq_kzg_prover.js:12-81), the Keccak
transcript schedule and the five fused device rounds behind the C ABI."""
import ctypes as C
import math
from collections import OrderedDict

from . import _lib
from ._lib import as_ptr
from .Keccak256Transcript import Keccak256Transcript
from .polynomial.evaluations import Evaluations
from .ptau_utils import readPTauHeader


def _host_column(ev):
    """an Evaluations column as handed in by the caller, zero-copy when it already is a host buffer"""
    return ev.host_buffer()


def prove(kind, pTauFilename, evalsFs, evalsTs, evalsSelF=None, evalsSelT=None, device=0, logger=None, trace=None):
    hdr = readPTauHeader(pTauFilename, device)                  # prover.js:15-16
    curve, nBitsPTau = hdr["curve"], hdr["power"]
    lib = curve.lib
    gs = kind == _lib.KZG_GRANDSUM

    if not isinstance(evalsFs, (list, tuple)):                  # :22-27
        evalsFs = [evalsFs]
    if not isinstance(evalsTs, (list, tuple)):
        evalsTs = [evalsTs]
    if len(evalsFs) != len(evalsTs):                            # :30-32
        raise ValueError("The lengths of the two vector multisets must be the same.")
    nPols = len(evalsFs)
    if nPols == 0:                                              # :34-36
        raise ValueError("The number of multisets must be greater than 0.")
    for i in range(nPols):                                      # :39-45
        if evalsFs[i].length() != evalsTs[i].length():
            raise ValueError("The %d-th multiset buffers must have the same length." % i)
        if evalsFs[i].length() != evalsFs[0].length():
            raise ValueError("The multiset buffers must all have the same length.")
    # :48-68.  Selectors that are not provided mean "all ones"; the reference materialises them and then finds
    # out they are all ones -- the outcome (isSelected = false) is known without building 2 x 32n bytes.
    if evalsSelF is None and evalsSelT is None:
        isSelected = False
    else:
        if evalsSelF is None:
            evalsSelF = Evaluations.getOneEvals(evalsFs[0].length(), curve)
        if evalsSelT is None:
            evalsSelT = Evaluations.getOneEvals(evalsTs[0].length(), curve)
        if evalsSelF.length() != evalsSelT.length():            # :56-60
            raise ValueError("The selection buffers must have the same length.")
        if evalsSelF.length() != evalsFs[0].length():
            raise ValueError("The selection buffers must have the same length as the multiset buffers.")
        isSelected = True                                       # :63-68
        if evalsSelF.isAllOnes() and evalsSelT.isAllOnes():
            isSelected = False
        elif evalsSelF.isAllZeros() and evalsSelT.isAllZeros():
            if logger:
                logger.warning("The selection buffers are all zeros. The argument is trivially satisfied.")
    length = evalsFs[0].length()
    nBits = math.ceil(math.log2(length)) if length > 0 else 0   # :70-71
    domainSize = 2 ** nBits
    if length != domainSize:                                    # :74-76
        raise ValueError("Polynomial length must be a power of two.")
    if nBitsPTau < nBits:                                       # :79-81
        raise ValueError("The Powers of Tau file is not sufficiently large to commit the polynomials.")

    srs, _ = curve.load_srs(pTauFilename, domainSize * 2)       # :83-85 (2n points, device-resident)
    isVector = nPols > 1
    acc = "S" if gs else "Z"

    prover = C.c_void_p()
    curve.check(lib.kzg_prover_create(curve.ctx, srs, kind, nBits, nPols, 1 if isSelected else 0, C.byref(prover)))
    try:
        proof = {"evaluations": OrderedDict(), "commitments": OrderedDict()}
        Cm, Ev = proof["commitments"], proof["evaluations"]
        transcript = Keccak256Transcript(curve)
        challenges = {}

        # ---- round 1: witness polynomials and their commitments (:144-179)
        cols_f = [_host_column(e) for e in evalsFs]
        cols_t = [_host_column(e) for e in evalsTs]
        pf = (C.c_void_p * nPols)(*[C.cast(as_ptr(c), C.c_void_p) for c in cols_f])
        pt = (C.c_void_p * nPols)(*[C.cast(as_ptr(c), C.c_void_p) for c in cols_t])
        sel_f = _host_column(evalsSelF) if isSelected else None
        sel_t = _host_column(evalsSelT) if isSelected else None
        n1 = 2 * nPols + (2 if isSelected else 0)
        out1 = bytearray(64 * n1)
        curve.check(lib.kzg_prover_round1(prover, pf, pt, as_ptr(sel_f), as_ptr(sel_t), as_ptr(out1)))
        for i in range(nPols):
            Cm["F%d" % i if isVector else "F"] = bytes(out1[128 * i:128 * i + 64])
            Cm["T%d" % i if isVector else "T"] = bytes(out1[128 * i + 64:128 * i + 128])
        if isSelected:
            Cm["selF"] = bytes(out1[128 * nPols:128 * nPols + 64])
            Cm["selT"] = bytes(out1[128 * nPols + 64:128 * nPols + 128])

        # ---- round 2: the grand-sum / grand-product polynomial (:181-231)
        for i in range(nPols):
            transcript.addPolCommitment(Cm["F%d" % i if isVector else "F"])
            transcript.addPolCommitment(Cm["T%d" % i if isVector else "T"])
        if isSelected:
            transcript.addPolCommitment(Cm["selF"])
            transcript.addPolCommitment(Cm["selT"])
        beta = None
        if isVector:
            beta = challenges["beta"] = transcript.getChallenge()
            transcript.addFieldElement(beta)
        gamma = challenges["gamma"] = transcript.getChallenge()
        out2 = bytearray(64)
        curve.check(lib.kzg_prover_round2(prover, as_ptr(beta), as_ptr(gamma), as_ptr(out2)))
        Cm[acc] = bytes(out2)

        # ---- round 3: the quotient polynomial (:233-286)
        transcript.addFieldElement(gamma)
        transcript.addPolCommitment(Cm[acc])
        alpha = challenges["alpha"] = transcript.getChallenge()
        out3 = bytearray(64)
        curve.check(lib.kzg_prover_round3(prover, as_ptr(alpha), as_ptr(out3)))
        Cm["Q"] = bytes(out3)

        # ---- round 4: evaluations (:288-318)
        transcript.addFieldElement(alpha)
        transcript.addPolCommitment(Cm["Q"])
        xi = challenges["xi"] = transcript.getChallenge()
        ne = int(lib.kzg_prover_n_evals(prover))
        out4 = bytearray(32 * ne)
        curve.check(lib.kzg_prover_round4(prover, as_ptr(xi), as_ptr(out4)))
        vals = [bytes(out4[32 * i:32 * i + 32]) for i in range(ne)]
        pos = 0
        for i in range(nPols):
            Ev["f%dxi" % i if isVector else "fxi"] = vals[pos]
            pos += 1
            if gs:
                Ev["t%dxi" % i if isVector else "txi"] = vals[pos]
                pos += 1
        if isSelected:
            Ev["selFxi"] = vals[pos]
            Ev["selTxi"] = vals[pos + 1]
            pos += 2
        Ev["sxiw" if gs else "zxiw"] = vals[pos]

        # ---- round 5: opening proofs (:320-413)
        transcript.addFieldElement(xi)
        for val in Ev.values():
            transcript.addFieldElement(val)
        v = challenges["v"] = transcript.getChallenge()
        out5 = bytearray(128)
        curve.check(lib.kzg_prover_round5(prover, as_ptr(v), as_ptr(out5)))
        Cm["Wxi"] = bytes(out5[:64])
        Cm["Wxiw"] = bytes(out5[64:])
        # side effect of the reference (:147-148): the callers' F / T evaluations are now in Montgomery form
        for i in range(nPols):
            for which, ev in ((0, evalsFs[i]), (1, evalsTs[i])):
                h = C.c_void_p()
                curve.check(lib.kzg_prover_take_evals(prover, i, which, C.byref(h)))
                ev.eval = curve.wrap(h)
        if trace is not None:
            trace["challenges"] = challenges
        return proof
    finally:
        lib.kzg_prover_destroy(prover)
