"""linpde_gp_b200 -- B200-native (sm_100a) implementation of linpde-gp's GP-PDE conditioning hot path.

Drop-in for ONE path of marvinpfoertner/linpde-gp (BASELINE.json:north_star): Gram / cross-covariance assembly of
operator-transformed kernels, FP64 Cholesky + triangular solves with a cached appendable factor, posterior mean /
covariance on test grids -- behind the reference's Python API (``linfuncops.diffops``, ``linfunctls``,
``randprocs.covfuncs``, ``GaussianProcess.condition_on_observations``).  The numerics are hand-written CUDA
kernels in ``liblpgp.so`` (C ABI, ``include/lpgp.h``) reached through ctypes; torch tensors are device buffers.
There is no CPU fallback: importing this package without the built library fails.
"""
from . import _lib  # noqa: F401  (fails loudly if liblpgp.so is missing)
from . import backend, domains, functions, linfuncops, linfunctls, linops, problems, randprocs, randvars
from .randprocs import ConditionalGaussianProcess, GaussianProcess

__version__ = "0.1.0"
