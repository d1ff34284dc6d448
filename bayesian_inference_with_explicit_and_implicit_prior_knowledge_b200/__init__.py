"""B200-native (sm_100a) particle-Gibbs-with-ancestor-sampling hot path of
VolkmannB/bayesian-inference-with-explicit-and-implicit-prior-knowledge.

Host code mirrors the reference's module layout (`PGAS`, `Filtering`, `BasisFunctions`,
`BayesianInferrence`) and calls hand-written CUDA through the C ABI of libpgas_b200.so
(include/pgas_b200.h).  There is no CPU fallback: every compute entry point raises without a
CUDA device.  The top-level `src/` package re-exports these modules under the reference's names.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "random", "models", "stats", "BasisFunctions", "Filtering", "BayesianInferrence", "PGAS", "distributed"]
