"""mset_eq_kzg_grandproduct_verifier -- drop-in for reference src/grandproduct/mset_eq_kzg_verifier.js:9-299 (host code)."""
from .._verifier_common import verify


def mset_eq_kzg_grandproduct_verifier(pTauFilename, proof, nBits, **kw):
    """-> bool; never raises on a bad proof"""
    return verify("gp", pTauFilename, proof, nBits, **kw)
