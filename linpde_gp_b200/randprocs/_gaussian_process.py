"""``GaussianProcess`` prior (pn/randprocs/_gaussian_process.py:15-79, _random_process.py:223-270) with the
``condition_on_observations`` entry point that linpde-gp patches in
(src/linpde_gp/randprocs/_gaussian_process/_conditional.py:402-406) and the ``L(gp)`` push-forwards of
src/linpde_gp/randprocs/_gaussian_process/_lintransforms.py:9-31."""
from __future__ import annotations

import numpy as np

from .. import functions, randvars
from . import covfuncs


class GaussianProcess:
    def __init__(self, mean, cov):
        if not isinstance(mean, functions.Function):
            raise TypeError("The mean function must have type `Function`.")
        if not isinstance(cov, covfuncs.CovarianceFunction):
            raise TypeError("The covariance function must have type `CovarianceFunction`.")
        if mean.input_shape != cov.input_shape:
            raise ValueError(
                f"The mean and covariance functions must have the same input shapes ({mean.input_shape} and "
                f"{cov.input_shape})."
            )
        if mean.output_shape != cov.output_shape_0 or mean.output_shape != cov.output_shape_1:
            raise ValueError(
                f"The output shapes of the mean function ({mean.output_shape}) and of the covariance function "
                f"({cov.output_shape_0}, {cov.output_shape_1}) must match."
            )
        if len(mean.output_shape) > 1:
            raise NotImplementedError("processes with more than one output axis are not on the accelerated path")
        self._mean = mean
        self._cov = cov

    @property
    def mean(self):
        return self._mean

    @property
    def cov(self):
        return self._cov

    @property
    def input_shape(self):
        return self._mean.input_shape

    @property
    def input_ndim(self):
        return self._mean.input_ndim

    @property
    def output_shape(self):
        return self._mean.output_shape

    @property
    def output_ndim(self):
        return len(self._mean.output_shape)

    def __call__(self, args) -> randvars.Normal:
        """Finite-dimensional marginal ``Normal(mean(x), cov.linop(x))`` -- the covariance stays a lazy,
        device-assembled operator, as in pn/randprocs/_gaussian_process.py:75-79."""
        x = np.asarray(args, dtype=np.double)
        return randvars.Normal(mean=np.asarray(self._mean(x), dtype=np.double).reshape(-1), cov=self._cov.linop(x))

    def var(self, args) -> np.ndarray:
        v = self._cov(np.asarray(args, dtype=np.double), None)
        if self.output_ndim:  # multi-output: the pointwise covariance is an (n, n) matrix per point
            v = np.diagonal(v, axis1=-2, axis2=-1)
        return v

    def std(self, args) -> np.ndarray:
        return np.sqrt(self.var(args))

    def condition_on_observations(self, Y, X=None, *, L=None, b=None):
        from ._conditional import ConditionalGaussianProcess

        return ConditionalGaussianProcess.from_observations(self, Y, X, L=L, b=b)


def apply_linfuncop_to_gp(L, gp: GaussianProcess) -> GaussianProcess:
    """``L(gp)`` for a LinearFunctionOperator (_lintransforms.py:25-31)."""
    mean = L(gp.mean)
    cov = L(L(gp.cov, argnum=1), argnum=0)
    return GaussianProcess(mean, cov)


def apply_linfunctl_to_gp(Lf, gp: GaussianProcess) -> randvars.Normal:
    """``L(gp)`` for a LinearFunctional: the (lazy) Gaussian ``Normal(L m, L k L^*)`` (_lintransforms.py:9-22)."""
    op, X = Lf._as_observation()  # pylint: disable=protected-access
    mean_fn = gp.mean if op is None else op(gp.mean)
    k = gp.cov if op is None else op(op(gp.cov, argnum=1), argnum=0)
    return randvars.Normal(mean=np.asarray(mean_fn(X)).reshape(-1), cov=k.linop(X))
