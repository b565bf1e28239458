from .polynomial import Polynomial
from .evaluations import Evaluations
from .polynomial_utils import computeZHEvaluation, computeL1Evaluation

__all__ = ["Polynomial", "Evaluations", "computeZHEvaluation", "computeL1Evaluation"]
