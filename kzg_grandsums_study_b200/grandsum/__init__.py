from .mset_eq_kzg_prover import mset_eq_kzg_grandsum_prover
from .mset_eq_kzg_verifier import mset_eq_kzg_grandsum_verifier, mset_eq_kzg_grandsum_verifier_batch
from .grandsum import ComputeSGrandSumPolynomial

__all__ = ["mset_eq_kzg_grandsum_prover", "mset_eq_kzg_grandsum_verifier", "mset_eq_kzg_grandsum_verifier_batch", "ComputeSGrandSumPolynomial"]
