"""mset_eq_kzg_grandsum_prover -- drop-in for reference src/grandsum/mset_eq_kzg_prover.js:12-435."""
from .. import _lib
from .._prover_common import prove


def mset_eq_kzg_grandsum_prover(pTauFilename, evalsFs, evalsTs, evalsSelF=None, evalsSelT=None, **kw):
    """-> proof = {evaluations: {...}, commitments: {...}} with the reference's keys and byte layouts
    (64 B affine Montgomery-LE commitments, 32 B Montgomery-LE evaluations, key insertion order kept)."""
    return prove(_lib.KZG_GRANDSUM, pTauFilename, evalsFs, evalsTs, evalsSelF, evalsSelT, **kw)
