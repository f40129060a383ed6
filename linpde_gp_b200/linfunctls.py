"""Linear functionals of the hot path (``linpde_gp.linfunctls``): point evaluation composed with an operator.

``_EvaluationFunctional`` (src/linpde_gp/linfunctls/_evaluation.py:10-60) evaluates a function at the points ``X``
(output layout: codomain_shape + batch_shape); ``CompositeLinearFunctional`` (``_arithmetic.py:92-140``) is
``linfunctl @ linop``, which is what ``LinearFunctionOperator.to_linfunctl(X)`` returns."""
from __future__ import annotations

import numpy as np

from .functions import _as_shape


class LinearFunctional:
    def __init__(self, input_shapes, output_shape):
        self._input_domain_shape = _as_shape(input_shapes[0])
        self._input_codomain_shape = _as_shape(input_shapes[1])
        self._output_shape = _as_shape(output_shape)

    @property
    def input_shapes(self):
        return (self._input_domain_shape, self._input_codomain_shape)

    @property
    def input_domain_shape(self):
        return self._input_domain_shape

    @property
    def input_codomain_shape(self):
        return self._input_codomain_shape

    @property
    def output_shape(self):
        return self._output_shape

    @property
    def output_size(self):
        return int(np.prod(self._output_shape)) if self._output_shape else 1

    # (operator or None, evaluation points) -- what the conditioning code needs from a functional
    def _as_observation(self):  # pragma: no cover - abstract
        raise NotImplementedError(f"{type(self).__name__} is not a point-evaluation observation")

    def __call__(self, f, /, **kwargs):
        from .randprocs import _conditional, _gaussian_process

        if isinstance(f, _conditional.ConditionalGaussianProcess):
            return f._apply_linfunctl(self)  # pylint: disable=protected-access
        if isinstance(f, _gaussian_process.GaussianProcess):
            return _gaussian_process.apply_linfunctl_to_gp(self, f)
        op, X = self._as_observation()
        g = f if op is None else op(f)
        return g(X)


class _EvaluationFunctional(LinearFunctional):
    def __init__(self, input_domain_shape, input_codomain_shape, X):
        from .randprocs import covfuncs

        # an intact TensorProductGrid is remembered: conditioning assembles its Gram blocks from Kronecker factors
        self._grid = X if covfuncs._grid_factors(X) is not None else None  # pylint: disable=protected-access
        X = np.asarray(X, dtype=np.double)
        input_domain_shape = _as_shape(input_domain_shape)
        input_codomain_shape = _as_shape(input_codomain_shape)
        nd = len(input_domain_shape)
        if X.shape[X.ndim - nd :] != input_domain_shape:
            raise ValueError(f"trailing shape of X {X.shape} must equal the input domain shape {input_domain_shape}")
        batch = X.shape[: X.ndim - nd]
        super().__init__((input_domain_shape, input_codomain_shape), input_codomain_shape + batch)
        self._X = X

    @property
    def X(self):
        return self._X

    def _as_observation(self):
        return None, (self._X if self._grid is None else self._grid)


class CompositeLinearFunctional(LinearFunctional):
    def __init__(self, *, linop, linfunctl):
        if tuple(linop.output_shapes[0]) != tuple(linfunctl.input_domain_shape):
            raise ValueError("shape mismatch between operator output and functional input")
        super().__init__(linop.input_shapes, linfunctl.output_shape)
        self._linop = linop
        self._linfunctl = linfunctl

    @property
    def linop(self):
        return self._linop

    @property
    def linfunctl(self):
        return self._linfunctl

    def _as_observation(self):
        inner_op, X = self._linfunctl._as_observation()
        if inner_op is not None:
            raise NotImplementedError("nested operator compositions are not supported")
        return self._linop, X
