"""mset_eq_kzg_grandproduct_prover -- drop-in for reference src/grandproduct/mset_eq_kzg_prover.js:12-415."""
from .. import _lib
from .._prover_common import prove


def mset_eq_kzg_grandproduct_prover(pTauFilename, evalsFs, evalsTs, evalsSelF=None, evalsSelT=None, **kw):
    """-> proof = {evaluations: {...}, commitments: {...}} (keys F,T[,selF,selT],Z,Q,Wxi,Wxiw / fxi[,selFxi,selTxi],zxiw)."""
    return prove(_lib.KZG_GRANDPRODUCT, pTauFilename, evalsFs, evalsTs, evalsSelF, evalsSelT, **kw)
