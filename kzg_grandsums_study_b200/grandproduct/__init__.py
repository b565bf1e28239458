from .mset_eq_kzg_prover import mset_eq_kzg_grandproduct_prover
from .mset_eq_kzg_verifier import mset_eq_kzg_grandproduct_verifier
from .grandproduct import ComputeZGrandProductPolynomial

__all__ = ["mset_eq_kzg_grandproduct_prover", "mset_eq_kzg_grandproduct_verifier", "ComputeZGrandProductPolynomial"]
