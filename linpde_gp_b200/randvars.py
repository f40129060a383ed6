"""Minimal random-variable containers (``pn.randvars.Normal/Constant``) used for observation noise ``b`` and for
the finite-dimensional marginals returned by ``GaussianProcess.__call__``."""
from __future__ import annotations

from typing import Optional

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


class Covariance:
    """Covariance between two random variables of shapes ``shape0`` and ``shape1``
    (src/linpde_gp/randvars/_covariance.py:13-134): simultaneously an array of shape ``shape0 + shape1`` and a
    ``(size0, size1)`` matrix / linear operator over the C-order flattened variables."""

    def __init__(self, shape0, shape1):
        self._shape0 = tuple(int(s) for s in np.atleast_1d(shape0)) if np.ndim(shape0) or shape0 != () else ()
        self._shape1 = tuple(int(s) for s in np.atleast_1d(shape1)) if np.ndim(shape1) or shape1 != () else ()

    shape0 = property(lambda self: self._shape0)
    shape1 = property(lambda self: self._shape1)
    ndim0 = property(lambda self: len(self._shape0))
    ndim1 = property(lambda self: len(self._shape1))
    size0 = property(lambda self: int(np.prod(self._shape0)) if self._shape0 else 1)
    size1 = property(lambda self: int(np.prod(self._shape1)) if self._shape1 else 1)

    @property
    def array(self) -> np.ndarray:  # pragma: no cover - abstract
        raise NotImplementedError

    @property
    def linop(self):  # pragma: no cover - abstract
        raise NotImplementedError

    @property
    def matrix(self) -> np.ndarray:
        return self.linop.todense()

    def flatten0(self, event0) -> np.ndarray:
        event0 = np.asarray(event0)
        if event0.shape != self._shape0:
            raise ValueError(f"The shape of the event must be the same as `shape0`, but {event0.shape} != {self._shape0}.")
        return event0.reshape(-1)

    def flatten1(self, event1) -> np.ndarray:
        event1 = np.asarray(event1)
        if event1.shape != self._shape1:
            raise ValueError(f"The shape of the event must be the same as `shape1`, but {event1.shape} != {self._shape1}.")
        return event1.reshape(-1)

    def unflatten0(self, vec0) -> np.ndarray:
        return np.asarray(vec0).reshape(self._shape0)

    def unflatten1(self, vec1) -> np.ndarray:
        return np.asarray(vec1).reshape(self._shape1)

    @property
    def T(self) -> "Covariance":
        return LinearOperatorCovariance(self.linop.T, shape0=self._shape1, shape1=self._shape0)

    def __neg__(self):
        return -1.0 * self

    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return LinearOperatorCovariance(float(other) * self.linop, shape0=self._shape0, shape1=self._shape1)
        return NotImplemented

    def __add__(self, other):
        if isinstance(other, Covariance):
            if other.shape0 != self._shape0 or other.shape1 != self._shape1:
                raise ValueError("shape mismatch")
            return LinearOperatorCovariance(self.linop + other.linop, shape0=self._shape0, shape1=self._shape1)
        return NotImplemented


class ArrayCovariance(Covariance):
    """Covariance given as a host array of shape ``shape0 + shape1`` (_covariance.py:137-194)."""

    def __init__(self, cov_array, shape0, shape1):
        super().__init__(shape0, shape1)
        self._array = np.asarray(cov_array, dtype=np.double)
        if self._array.shape != self.shape0 + self.shape1:
            raise ValueError(f"`cov_array` must have shape {self.shape0 + self.shape1}, got {self._array.shape}")
        self._linop = None

    @classmethod
    def from_scalar(cls, cov):
        return cls(np.asarray(cov, dtype=np.double).reshape(()), (), ())

    @property
    def array(self):
        return self._array

    @property
    def matrix(self):
        return self._array.reshape(self.size0, self.size1)

    @property
    def linop(self):
        if self._linop is None:
            self._linop = linops.Matrix(self.matrix)
        return self._linop


class LinearOperatorCovariance(Covariance):
    """Covariance given as a (lazy, device-resident) ``(size0, size1)`` linear operator (_covariance.py:197-230)."""

    def __init__(self, cov_linop, shape0, shape1):
        super().__init__(shape0, shape1)
        if tuple(cov_linop.shape) != (self.size0, self.size1):
            raise ValueError(f"`cov_linop` must have shape {(self.size0, self.size1)}, got {tuple(cov_linop.shape)}")
        self._linop = cov_linop

    @property
    def linop(self):
        return self._linop

    @property
    def array(self):
        return self.matrix.reshape(self.shape0 + self.shape1)


def asrandvar(b):
    """``pn.randvars.asrandvar``: random variables pass through, anything array-like (lists, ndarrays, numpy
    scalars -- which do have a ``.mean`` METHOD) becomes a ``Constant``."""
    if isinstance(b, (Normal, Constant)):
        return b
    return Constant(np.asarray(b, dtype=np.double))


def condition_normal_on_observations(prior: "Normal", observations, noise: "Optional[Normal]", transform=None) -> "Normal":
    r"""Conditions a Gaussian random vector on linearly transformed observations ``y = A x + eps`` with
    ``x ~ N(mu, Sigma)``, ``eps ~ N(b, Lambda)`` (src/linpde_gp/randvars/_normal.py:8-71; the finite-dimensional analogue
    of ``condition_on_observations``; also installed as ``Normal.condition_on_observations`` like the reference does).

    The Gram matrix ``A Sigma A^T + Lambda`` is factored by the device Cholesky and the gain computed by a multi-right-hand
    side ``potrs``; the matrix products run through the DMMA GEMM (``linops.Matrix @``)."""
    y = np.asarray(observations, dtype=np.double)
    mu = prior.mean.reshape(-1)
    Sigma = np.asarray(prior.dense_cov, dtype=np.double)
    A = transform
    if A is not None:
        A = A.todense() if isinstance(A, linops.LinearOperator) else np.asarray(A, dtype=np.double)
        if A.ndim == 1:  # one scalar observation (_normal.py:26-29)
            A = A[None, :]
            y = y.reshape(1)
            if noise is not None:
                noise = Normal(np.reshape(noise.mean, (1,)), np.reshape(noise.dense_cov, (1, 1)))
        if A.ndim != 2 or A.shape[1] != mu.size:
            raise ValueError(f"`transform` must have shape (m, {mu.size}), got {A.shape}")
    y = y.reshape(-1)
    if A is None:
        cross, pred_mean, pred_cov = Sigma, mu, Sigma  # Cov(y, x)
    else:
        cross = linops.Matrix(A) @ Sigma
        pred_mean, pred_cov = A @ mu, linops.Matrix(cross) @ A.T
    if noise is not None:
        if noise.mean.size != pred_mean.size:
            raise ValueError("`noise` must have the shape of the observations")
        pred_mean = pred_mean + noise.mean.reshape(-1)
        pred_cov = pred_cov + np.asarray(noise.dense_cov, dtype=np.double)
    if y.size != pred_mean.size:
        raise ValueError(f"{pred_mean.size} observations expected, got {y.size}")
    gram = linops.Matrix(0.5 * (pred_cov + pred_cov.T))
    gram.is_symmetric = True
    gram.is_positive_definite = True
    gain_t = gram.solve(cross)  # (m, n): G^{-1} Cov(y, x)
    mean = mu + gain_t.T @ (y - pred_mean)
    cov = Sigma - linops.Matrix(np.ascontiguousarray(cross.T)) @ gain_t
    return Normal(mean.reshape(prior.mean.shape), 0.5 * (cov + cov.T))


Normal.condition_on_observations = condition_normal_on_observations
