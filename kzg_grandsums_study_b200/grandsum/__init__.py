from .mset_eq_kzg_prover import mset_eq_kzg_grandsum_prover
from .grandsum import ComputeSGrandSumPolynomial

__all__ = ["mset_eq_kzg_grandsum_prover", "ComputeSGrandSumPolynomial"]
