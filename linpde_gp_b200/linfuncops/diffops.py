"""Linear differential operators (symbolic layer) -- the reference's ``linpde_gp.linfuncops.diffops`` API for the
operators on the hot path: same names, constructor arguments and coefficient representation
(src/linpde_gp/linfuncops/diffops/{_coefficients,_lindiffop,_partial_derivative,_directional_derivative,
_laplacian,_heat,_arithmetic}.py).  Applying an operator to a covariance function returns a *transformed*
covariance function whose numerics run on the GPU (``randprocs/covfuncs.py``)."""
from __future__ import annotations

from collections.abc import Mapping
from copy import deepcopy

import numpy as np

from ..functions import _as_shape
from ._linfuncop import LinearFunctionOperator, SumLinearFunctionOperator


class MultiIndex:
    """Multi-index of a partial derivative (``_coefficients.py:9-62``)."""

    def __init__(self, multi_index):
        self._multi_index = np.asarray(multi_index, dtype=int)
        if np.any(self._multi_index < 0):
            raise ValueError(f"Multi-index {multi_index} contains negative entries.")
        self._multi_index.setflags(write=False)

    @classmethod
    def from_index(cls, index, shape, order):
        mi = np.zeros(shape, dtype=int)
        mi[index] = order
        return cls(mi)

    @property
    def order(self) -> int:
        return int(np.sum(self._multi_index))

    @property
    def is_mixed(self) -> bool:
        return bool(np.count_nonzero(self._multi_index) > 1)

    @property
    def array(self):
        return self._multi_index

    @property
    def shape(self):
        return self._multi_index.shape

    def __getitem__(self, index):
        return self._multi_index[index]

    def as_tuple(self):
        return tuple(int(v) for v in np.atleast_1d(self._multi_index).reshape(-1))

    def __hash__(self):
        return hash(self._multi_index.tobytes())

    def __eq__(self, other):
        if not isinstance(other, MultiIndex):
            return NotImplemented
        return self.shape == other.shape and bool(np.all(self.array == other.array))

    def __repr__(self):
        return f"MultiIndex({self._multi_index.tolist()})"


class PartialDerivativeCoefficients(Mapping):
    """``{codomain_index: {MultiIndex: coefficient}}`` (``_coefficients.py:65-200``)."""

    def __init__(self, coefficient_dict, input_domain_shape, input_codomain_shape):
        input_domain_shape = _as_shape(input_domain_shape)
        input_codomain_shape = _as_shape(input_codomain_shape)
        n = 0
        for codomain_idx, inner in coefficient_dict.items():
            if len(codomain_idx) != len(input_codomain_shape) or not all(
                x < y for x, y in zip(codomain_idx, input_codomain_shape)
            ):
                raise ValueError(f"Codomain index {codomain_idx} does not match shape {input_codomain_shape}.")
            for mi in inner:
                if mi.shape != input_domain_shape:
                    raise ValueError(
                        f"Multi-index shape {mi.shape} does not match input domain shape {input_domain_shape}."
                    )
                n += 1
        self._dict = coefficient_dict
        self._num_entries = n
        self._input_domain_shape = input_domain_shape
        self._input_codomain_shape = input_codomain_shape

    @property
    def num_entries(self):
        return self._num_entries

    @property
    def has_mixed(self):
        return any(mi.is_mixed for inner in self._dict.values() for mi in inner)

    @property
    def input_domain_shape(self):
        return self._input_domain_shape

    @property
    def input_codomain_shape(self):
        return self._input_codomain_shape

    def __getitem__(self, codomain_idx):
        return self._dict[codomain_idx]

    def __len__(self):
        return len(self._dict)

    def __iter__(self):
        return iter(self._dict)

    def __neg__(self):
        return -1.0 * self

    def __add__(self, other):
        if not isinstance(other, PartialDerivativeCoefficients):
            return NotImplemented
        if self.input_domain_shape != other.input_domain_shape:
            raise ValueError("Cannot add coefficients with different input domain shapes")
        if self.input_codomain_shape != other.input_codomain_shape:
            raise ValueError("Cannot add coefficients with different input codomain shapes")
        new = deepcopy(self._dict)
        for cidx, inner in other.items():
            tgt = new.setdefault(cidx, {})
            for mi, c in inner.items():
                tgt[mi] = tgt.get(mi, 0.0) + c
        return PartialDerivativeCoefficients(new, self.input_domain_shape, self.input_codomain_shape)

    def __sub__(self, other):
        return self + (-other)

    def __rmul__(self, other):
        if np.ndim(other) != 0:
            return NotImplemented
        scaled = {cidx: {mi: float(other) * c for mi, c in inner.items()} for cidx, inner in self._dict.items()}
        return PartialDerivativeCoefficients(scaled, self.input_domain_shape, self.input_codomain_shape)


class LinearDifferentialOperator(LinearFunctionOperator):
    """Scalar-output linear differential operator (``_lindiffop.py:24-160``)."""

    def __init__(self, coefficients: PartialDerivativeCoefficients, input_shapes):
        input_shapes = (_as_shape(input_shapes[0]), _as_shape(input_shapes[1]))
        if coefficients.input_domain_shape != input_shapes[0]:
            raise ValueError()
        if coefficients.input_codomain_shape != input_shapes[1]:
            raise ValueError()
        super().__init__(input_shapes=input_shapes, output_shapes=(input_shapes[0], ()))
        self._coefficients = coefficients

    @property
    def coefficients(self):
        return self._coefficients

    @property
    def has_mixed(self):
        return self._coefficients.has_mixed

    @property
    def order(self) -> int:
        return max((mi.order for mi in self._coefficients[()]), default=0)

    def _terms(self):
        """Flat ``{multi_index_tuple: coeff}`` of the scalar-output operator."""
        out = {}
        for mi, c in self._coefficients[()].items():
            key = mi.as_tuple()
            out[key] = out.get(key, 0.0) + float(c)
        return out

    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledLinearDifferentialOperator(self, scalar=other)
        return NotImplemented

    def __neg__(self):
        return -1.0 * self


class ScaledLinearDifferentialOperator(LinearDifferentialOperator):
    """``scalar * lindiffop`` (``diffops/_arithmetic.py:10-70``)."""

    def __init__(self, lindiffop, /, scalar):
        if np.ndim(scalar) != 0:
            raise ValueError()
        self._lindiffop = lindiffop
        self._scalar = np.asarray(scalar, dtype=np.double)
        super().__init__(
            coefficients=float(self._scalar) * lindiffop.coefficients,
            input_shapes=lindiffop.input_shapes,
        )

    @property
    def lindiffop(self):
        return self._lindiffop

    @property
    def scalar(self):
        return self._scalar

    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledLinearDifferentialOperator(self._lindiffop, scalar=np.asarray(other) * self._scalar)
        return NotImplemented

    def __repr__(self):
        return f"{self._scalar} * {self._lindiffop}"


class PartialDerivative(LinearDifferentialOperator):
    def __init__(self, multi_index: MultiIndex):
        super().__init__(
            coefficients=PartialDerivativeCoefficients({(): {multi_index: 1.0}}, multi_index.shape, ()),
            input_shapes=(multi_index.shape, ()),
        )
        self._multi_index = multi_index

    @property
    def multi_index(self):
        return self._multi_index

    @property
    def is_mixed(self):
        return self._multi_index.is_mixed

    def __repr__(self):
        return f"{self.__class__.__name__}(multi_index={self.multi_index})"


class TimeDerivative(PartialDerivative):
    def __init__(self, domain_shape):
        domain_shape = _as_shape(domain_shape)
        if len(domain_shape) == 0:
            mi = 1
        elif len(domain_shape) == 1:
            mi = (1,) + (0,) * (domain_shape[0] - 1)
        else:
            raise ValueError()
        super().__init__(MultiIndex(mi))


class Derivative(PartialDerivative):
    """n-th derivative of a univariate function (``diffops/_derivative.py``)."""

    def __init__(self, order: int):
        if order < 0:
            raise ValueError(f"Order must be >= 0, but got {order}.")
        super().__init__(MultiIndex(order))


class DirectionalDerivative(LinearDifferentialOperator):
    def __init__(self, direction):
        direction = np.asarray(direction, dtype=np.double)
        coeffs = PartialDerivativeCoefficients(
            {(): {MultiIndex.from_index(i, direction.shape, 1): float(c) for i, c in np.ndenumerate(direction) if c != 0.0}},
            input_domain_shape=direction.shape,
            input_codomain_shape=(),
        )
        super().__init__(coefficients=coeffs, input_shapes=(direction.shape, ()))
        self._direction = direction

    @property
    def direction(self):
        return self._direction


class WeightedLaplacian(LinearDifferentialOperator):
    def __init__(self, weights):
        weights = np.asarray(weights, dtype=np.double)
        coeffs = PartialDerivativeCoefficients(
            {(): {MultiIndex.from_index(i, weights.shape, 2): float(c) for i, c in np.ndenumerate(weights) if c != 0.0}},
            input_domain_shape=weights.shape,
            input_codomain_shape=(),
        )
        super().__init__(coefficients=coeffs, input_shapes=(weights.shape, ()))
        self._weights = weights

    @property
    def weights(self):
        return self._weights


class Laplacian(WeightedLaplacian):
    def __init__(self, domain_shape):
        super().__init__(np.ones(_as_shape(domain_shape), dtype=np.double))


class SpatialLaplacian(WeightedLaplacian):
    def __init__(self, domain_shape):
        domain_shape = _as_shape(domain_shape)
        if len(domain_shape) != 1 or domain_shape[0] < 2:
            raise ValueError()
        weights = np.ones(domain_shape, dtype=np.double)
        weights[0] = 0
        super().__init__(weights)


class HeatOperator(SumLinearFunctionOperator):
    """``d/dt - alpha * Laplacian_x`` as the sum of a ``TimeDerivative`` and a ``WeightedLaplacian``
    (``_heat.py:14-39``)."""

    def __init__(self, domain_shape, alpha=1.0):
        domain_shape = _as_shape(domain_shape)
        if len(domain_shape) != 1:
            raise ValueError("The `HeatOperator` only applies to functions with `input_ndim == 1`.")
        self._alpha = float(alpha)
        w = np.zeros(domain_shape, dtype=np.double)
        w[1:] = -self._alpha
        super().__init__(TimeDerivative(domain_shape), WeightedLaplacian(w))

    @property
    def alpha(self):
        return self._alpha
