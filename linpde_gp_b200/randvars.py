"""Minimal random-variable containers (``pn.randvars.Normal/Constant``) used for observation noise ``b`` and for
the finite-dimensional marginals returned by ``GaussianProcess.__call__``."""
from __future__ import annotations

import numpy as np

from . import linops


class Constant:
    def __init__(self, support):
        self._support = np.asarray(support, dtype=np.double)

    @property
    def support(self):
        return self._support

    mean = support

    @property
    def shape(self):
        return self._support.shape

    @property
    def cov(self):
        n = self._support.size
        return linops.Scaling(np.zeros(n))


class Normal:
    """``Normal(mean, cov)``; ``cov`` is a (N, N) array, a 1-D array of variances is NOT accepted (as in probnum) --
    pass ``linops.Scaling(variances)`` for diagonal noise."""

    def __init__(self, mean, cov):
        self._mean = np.asarray(mean, dtype=np.double)
        if isinstance(cov, linops.LinearOperator):
            self._cov = cov
        else:
            cov = np.asarray(cov, dtype=np.double)
            if cov.ndim == 0 and self._mean.ndim == 0:
                cov = cov.reshape(1, 1)
            n = self._mean.size
            if cov.shape != (n, n):
                raise ValueError(f"The covariance matrix must have shape ({n}, {n}), got {cov.shape}.")
            self._cov = cov
        n = self._mean.size
        if tuple(self._cov.shape) != (n, n):
            raise ValueError(f"The covariance must have shape ({n}, {n}), got {tuple(self._cov.shape)}.")

    @property
    def mean(self):
        return self._mean

    @property
    def cov(self):
        return self._cov

    @property
    def shape(self):
        return self._mean.shape

    def __neg__(self):
        """``-b``: same covariance (kept lazy), negated mean -- e.g. ``b=-f_prior(X)`` for an uncertain right-hand
        side (experiments/0003_poisson_1d_inverse_rhs.ipynb cell 19)."""
        return Normal(-self._mean, self._cov)

    @property
    def dense_cov(self):
        return self._cov.todense() if isinstance(self._cov, linops.LinearOperator) else self._cov

    @property
    def var(self):
        return np.diag(self.dense_cov).reshape(self._mean.shape)

    @property
    def std(self):
        return np.sqrt(self.var)


def asrandvar(b):
    if isinstance(b, (Normal, Constant)):
        return b
    if np.ndim(b) >= 0 and not hasattr(b, "mean"):
        return Constant(b)
    return b
