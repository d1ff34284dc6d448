"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy float64 restatement of the reference's PGAS hot path
(VolkmannB/bayesian-inference-with-explicit-and-implicit-prior-knowledge).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package; the product path (the
``bayesian_inference_with_explicit_and_implicit_prior_knowledge_b200`` package
and the ``src`` shim) never does.

PARITY STATUS: pinned against the reference's own SOURCE, not against JAX's
numerics.  The reference runs on jax 0.4.38 / jaxlib 0.4.38 / equinox 0.12.2
(pyproject.toml:7-12), none of which is installable here; it ships no tests or
golden vectors and plots/*.mat are Git-LFS stubs.  tests/golden/jaxshim/ is a
NumPy/SciPy-backed stand-in for the ~60 jax / equinox entry points the
reference calls; with it, tests/golden/make_reference_golden.py imports the
UNMODIFIED reference modules (src/Filtering.py, BasisFunctions.py,
BayesianInferrence.py, StateSpaceModel.py, PGAS.py, Algorithm1/2/3.py,
SingleMassOscillator.py, Vehicle.py) from /root/reference, runs them on small
seeded problems and records their outputs together with every random variate
they consumed (tests/golden/reference_golden.npz).  tests/test_reference_pins.py
feeds those variates to this restatement and requires the reference's outputs
back (indices exactly, float64 within 1e-9 relative); tests/
test_gpu_reference_pins.py does the same for the CUDA path.  What stays
unpinned is XLA-CPU's floating-point behaviour itself (the stand-in computes
with NumPy/LAPACK) and JAX's threefry stream.  The restatement still follows
the reference line by line (each function cites file:line) and is additionally
checked against analytic identities and SciPy (tests/test_oracle_pins.py).

Random numbers: JAX's threefry stream cannot be reproduced without JAX, so every
function takes *injected variates* in place of a key.
"""
