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
installed (no network), so the reference cannot be imported or executed here, and it
ships no golden vectors for the EK1 path (``tests/test_pdefilter.py:143-146`` only checks
for NaNs).  Consequently:

* the square-root primitives and the IWP prior ARE pinned: the reference's own
  known-answer tests (``tests/test_base/test_sqrt.py:37-109``,
  ``tests/test_base/test_iwp.py:20-62``, ``tests/test_discretize.py:64-71``,
  ``tests/test_odetools/test_step.py:28-45``) are re-run against this oracle in
  ``tests/test_oracle_*.py``;
* EK1 step / trajectory VALUES are **parity unpinned** -- the oracle follows the
  reference source line by line and uses the same LAPACK routines that jaxlib-CPU
  dispatches to (``dgeqrf``/``dtrsm``/``dgetrf``/``dpotrf`` through NumPy/SciPy), but no
  output of the real reference could be generated to anchor it.  ``tests/golden/`` holds
  vectors produced by THIS oracle (script: ``tests/golden/make_golden.py``) so that later
  changes to oracle or kernels are detected; they are regression anchors, not reference
  outputs.
"""
