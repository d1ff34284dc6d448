"""Reference-named shim: `src.Algorithm2` of the reference maps to the B200 implementation."""
from bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200.Algorithm2 import *  # noqa: F401,F403
from bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200 import Algorithm2 as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
