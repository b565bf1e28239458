"""mset_eq_kzg_grandproduct_verifier -- drop-in for reference src/grandproduct/mset_eq_kzg_verifier.js:9-299 (host code)."""
from .._verifier_common import verify, verify_batch


def mset_eq_kzg_grandproduct_verifier(pTauFilename, proof, nBits, **kw):
    """-> bool; never raises on a bad proof"""
    return verify("gp", pTauFilename, proof, nBits, **kw)


def mset_eq_kzg_grandproduct_verifier_batch(pTauFilename, proofs, nBits, **kw):
    """-> bool: True iff every proof of the list verifies.  One device MSM per side over all the proofs' commitments
    and ONE pairing product instead of ~10 scalar multiplications and a pairing product per proof"""
    return verify_batch("gp", pTauFilename, proofs, nBits, **kw)
