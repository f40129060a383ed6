"""CPU oracle for the GP-PDE conditioning hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy/scipy restatement of the reference's algorithm (marvinpfoertner/linpde-gp @ f1fc705,
vendored probnum fork @ 67d7d43) for the one path this repository accelerates: Gram / cross-covariance
assembly of operator-transformed kernels, FP64 Cholesky + triangular solves with a bordered (appendable)
factor, and posterior mean / covariance evaluation.  Every function cites the reference file:line it
follows (paths relative to /root/reference; ``pn`` = ``probnum/src/probnum``).

Who may use this package: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs -- and there only as the checker / the timed CPU baseline.  The product package
``linpde_gp_b200`` never imports it and has no CPU fallback.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the real reference in the build container through
``oracle/refshim.py`` and freezes its outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks this restatement against those vectors (and against the reference's own doctest known answers),
so the oracle is anchored on outputs of the reference itself.
"""

from . import covfuncs, gp, linalg  # noqa: F401
