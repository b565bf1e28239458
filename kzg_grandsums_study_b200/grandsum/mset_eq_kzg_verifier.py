"""mset_eq_kzg_grandsum_verifier -- drop-in for reference src/grandsum/mset_eq_kzg_verifier.js:9-313 (host code)."""
from .._verifier_common import verify


def mset_eq_kzg_grandsum_verifier(pTauFilename, proof, nBits, **kw):
    """-> bool; never raises on a bad proof (verifier.js:50,61,184-192)"""
    return verify("gs", pTauFilename, proof, nBits, **kw)
