from .mset_eq_kzg_prover import mset_eq_kzg_grandproduct_prover
from .grandproduct import ComputeZGrandProductPolynomial

__all__ = ["mset_eq_kzg_grandproduct_prover", "ComputeZGrandProductPolynomial"]
