"""CPU oracle for the pnmol EK1 hot path  --  TEST INFRASTRUCTURE ONLY.

This package is a float64 NumPy/SciPy restatement of the reference algorithm
(schmidtjonathan/pnmol-experiments, ``src/pnmol``; every function cites the
reference file:line it follows).  It exists so that the CUDA path can be checked
against the reference's arithmetic; it is NOT part of the product:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
  ``--impl reference`` legs may import it -- always as the checker or as the reported
  CPU baseline, never as a fallback for the CUDA path;
* the product package (``pnmol-experiments_b200/pnmol_b200``) never imports it and
  raises when its CUDA library is missing.

PARITY PINNING STATUS
---------------------
The reference is pure Python on top of JAX (jax/jaxlib <= 0.3.1, un-pinned in
``setup.cfg:16-17``) + tornadox.  Neither is installed in this image and nothing can be
installed (no network), and the reference ships no golden vectors for the EK1 path
(``tests/test_pdefilter.py:143-146`` only checks for NaNs).  The oracle is pinned in two ways:

* the square-root primitives, the IWP prior, the FD weights and the step rules are pinned to
  the reference's own known-answer tests (``tests/test_base/test_sqrt.py:37-109``,
  ``tests/test_base/test_iwp.py:20-62``, ``tests/test_discretize.py:64-71``,
  ``tests/test_odetools/test_step.py:28-45``), re-run against this oracle in
  ``tests/test_oracle_primitives.py``;
* EK1 step / trajectory VALUES are pinned to outputs of **the reference's own source**:
  ``tests/golden/make_reference_golden.py`` executes the unmodified reference files
  (``src/pnmol/white.py``, ``latent.py``, ``pdefilter.py``, ``base/{sqrt,iwp,stacked_ssm,rv}.py``,
  ``odetools/step.py``) for all four solver classes -- ``solve``, ``solution_generator``,
  ``simulate_final_state`` with ``Constant`` and ``Adaptive`` steps -- on a NumPy stand-in for the
  used slice of the JAX API (``tests/golden/jax_numpy_shim``: ``jax.numpy`` -> NumPy,
  ``jax.scipy.linalg`` -> SciPy, ``jit`` -> identity), i.e. on the same LAPACK routines that
  jaxlib's CPU backend dispatches to.  ``tests/golden/reference_*.npz`` are those outputs;
  ``tests/test_reference_golden.py`` checks the oracle against them (equal to rounding: means
  1e-12, covariances 1e-11, all by-products and info counters) and the CUDA path at the
  north-star tolerances.  What this does NOT cover: XLA's own float64 code generation (fusion,
  FMA contraction) of the real jaxlib, which can differ from NumPy at the 1e-16 level; and the
  problem set-up (discretisation, kernels), which needs JAX autodiff and is supplied by
  ``oracle/setup_np.py`` (pinned to the reference's ``test_discretize`` known answers).
  ``tests/golden/<config>.npz`` (without the ``reference_`` prefix) are older regression anchors
  produced by this oracle itself (``tests/golden/make_golden.py``).
"""
