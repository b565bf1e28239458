"""kzg_grandsums_study_b200 -- B200-native (sm_100a) backend for the prover hot path of
xavi-pinsach/kzg-grandsums-study.  Host layer = the reference's own module layout; compute = libkzgb200.so."""
from .curve import getCurveFromName, getCurveFromQ, Curve, DeviceBuffer  # noqa: F401
from ._lib import KzgError  # noqa: F401
