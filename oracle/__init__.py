"""CPU oracle for the GridNet hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``gridnext_b200``) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the timed CPU arm.

What it restates (reference paths relative to /root/reference):

* ``hexagdly.Conv2d`` (third-party PyPI package "HexagDLy", un-vendored and un-pinned:
  requirements.txt:11, imported at gridnext/gridnet_models.py:10).  The package is
  absent from the reference tree and cannot be installed offline, so its published
  algorithm is restated in ``oracle/hexconv_ref.py`` three independent ways.
  **PARITY UNPINNED for this dependency**: the reference ships no test, golden vector or
  checkpoint at that boundary.
* ``gridnext/gridnet_models.py`` (GridNet / GridNetHex / GridNetHexOddr / GridNetHexMM),
  ``gridnext/densenet.py``, the tutorial count MLP, ``gridnext/training.py:141-171`` and
  ``gridnext/imgprocess.py:162-238`` -- restated functionally in ``oracle/gridnet_ref.py``
  and ``oracle/gather_ref.py``.  These ARE pinned: ``oracle/make_golden.py`` imports the
  real reference modules in the build container (with the hexagdly shim and a matplotlib
  stub), runs them on seeded inputs and commits the outputs under ``tests/golden/``;
  ``tests/test_oracle_golden.py`` checks the restatement against those vectors.
"""
