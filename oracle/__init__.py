"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy float64 restatement of the reference's PGAS hot path
(VolkmannB/bayesian-inference-with-explicit-and-implicit-prior-knowledge).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; the product path (the
``bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200`` package
and the ``src`` shim) never does.

PARITY UNPINNED: the reference runs on jax 0.4.38 / jaxlib 0.4.38 /
equinox 0.12.2 (pyproject.toml:7-12), none of which is installable here, it
ships no tests / golden vectors, and plots/*.mat are Git-LFS stubs.  The
restatement therefore follows the reference source line by line (each function
cites file:line) and encodes the published JAX semantics of the primitives it
calls (softmax, searchsorted side='left', clamped gathers, chisquare =
2*Gamma(df/2), multivariate_normal = mean + chol(cov) z); it is pinned only
against analytic identities and SciPy cross-checks (tests/test_oracle_*.py),
not against outputs of the reference itself.

Random numbers: JAX's threefry stream cannot be reproduced without JAX, so every
function takes *injected variates* in place of a key.
"""
