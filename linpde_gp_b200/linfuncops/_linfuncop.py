"""``LinearFunctionOperator`` and its arithmetic (src/linpde_gp/linfuncops/_linfuncop.py:16-105,
_arithmetic.py:12-160).  Operators are symbolic; calling one dispatches on the argument type."""
from __future__ import annotations

import functools
import operator

import numpy as np

from ..functions import Constant, Function, Zero, _as_shape


class LinearFunctionOperator:
    def __init__(self, input_shapes, output_shapes):
        self._input_domain_shape = _as_shape(input_shapes[0])
        self._input_codomain_shape = _as_shape(input_shapes[1])
        self._output_domain_shape = _as_shape(output_shapes[0])
        self._output_codomain_shape = _as_shape(output_shapes[1])

    @property
    def input_shapes(self):
        return (self._input_domain_shape, self._input_codomain_shape)

    @property
    def output_shapes(self):
        return (self._output_domain_shape, self._output_codomain_shape)

    @property
    def input_domain_shape(self):
        return self._input_domain_shape

    @property
    def input_codomain_shape(self):
        return self._input_codomain_shape

    @property
    def output_domain_shape(self):
        return self._output_domain_shape

    @property
    def output_codomain_shape(self):
        return self._output_codomain_shape

    # -- flat representation used by the device lowering -----------------------------------------------
    def _terms(self):  # pragma: no cover - abstract
        raise NotImplementedError(f"{type(self).__name__} has no partial-derivative representation")

    # -- application --------------------------------------------------------------------------------------
    def __call__(self, f, /, **kwargs):
        from ..randprocs import _conditional, _gaussian_process, covfuncs

        if isinstance(f, covfuncs.CovarianceFunction):
            return covfuncs.apply_linfuncop(self, f, argnum=kwargs.get("argnum", 0))
        if isinstance(f, _conditional.ConditionalGaussianProcess):
            return f._apply_linfuncop(self)  # pylint: disable=protected-access
        if isinstance(f, _gaussian_process.GaussianProcess):
            return _gaussian_process.apply_linfuncop_to_gp(self, f)
        if isinstance(f, Function):
            return self._apply_to_function(f)
        raise NotImplementedError(
            f"{type(self).__name__} cannot be applied to {type(f).__name__} without the reference's autodiff fallback"
        )

    def _apply_to_function(self, f):
        terms = self._terms()
        order0 = sum(c for mi, c in terms.items() if sum(mi) == 0)
        if isinstance(f, Zero):
            return Zero(self.output_domain_shape, self.output_codomain_shape)
        if isinstance(f, Constant):
            return Constant(self.output_domain_shape, order0 * f.value)
        raise NotImplementedError("only Zero / Constant functions can be differentiated in closed form")

    def to_linfunctl(self, X):
        """``_EvaluationFunctional(X) @ self`` (``_linfuncop.py:93-105``)."""
        from .. import linfunctls

        return linfunctls.CompositeLinearFunctional(
            linop=self,
            linfunctl=linfunctls._EvaluationFunctional(  # pylint: disable=protected-access
                input_domain_shape=self.output_domain_shape,
                input_codomain_shape=self.output_codomain_shape,
                X=X,
            ),
        )

    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledLinearFunctionOperator(self, scalar=other)
        return NotImplemented

    def __neg__(self):
        return -1.0 * self

    def __add__(self, other):
        if isinstance(other, LinearFunctionOperator):
            return SumLinearFunctionOperator(self, other)
        return NotImplemented

    def __sub__(self, other):
        if isinstance(other, LinearFunctionOperator):
            return SumLinearFunctionOperator(self, -other)
        return NotImplemented


class ScaledLinearFunctionOperator(LinearFunctionOperator):
    def __init__(self, linfuncop, /, scalar):
        if np.ndim(scalar) != 0:
            raise ValueError()
        self._linfuncop = linfuncop
        self._scalar = np.asarray(scalar, dtype=np.double)
        super().__init__(input_shapes=linfuncop.input_shapes, output_shapes=linfuncop.output_shapes)

    @property
    def linfuncop(self):
        return self._linfuncop

    @property
    def scalar(self):
        return self._scalar

    def _terms(self):
        return {mi: float(self._scalar) * c for mi, c in self._linfuncop._terms().items()}

    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledLinearFunctionOperator(self._linfuncop, scalar=np.asarray(other) * self._scalar)
        return NotImplemented

    def __repr__(self):
        return f"{self._scalar} * {self._linfuncop}"


class SumLinearFunctionOperator(LinearFunctionOperator):
    def __init__(self, *summands):
        self._summands = tuple(summands)
        first = self._summands[0]
        assert all(s.input_shapes == first.input_shapes and s.output_shapes == first.output_shapes for s in self._summands)
        super().__init__(input_shapes=first.input_shapes, output_shapes=first.output_shapes)

    @property
    def summands(self):
        return self._summands

    def _terms(self):
        out = {}
        for s in self._summands:
            for mi, c in s._terms().items():
                out[mi] = out.get(mi, 0.0) + c
        return out

    def __repr__(self):
        return " + ".join(str(s) for s in self._summands)


class Identity(LinearFunctionOperator):
    """Order-zero operator (the reference uses ``PartialDerivative(MultiIndex(zeros))``, diffops/_registry.py:41-50)."""

    def __init__(self, domain_shape, codomain_shape=()):
        super().__init__(input_shapes=(domain_shape, codomain_shape), output_shapes=(domain_shape, codomain_shape))

    def _terms(self):
        d = int(np.prod(self.input_domain_shape)) if self.input_domain_shape else 1
        return {tuple([0] * d): 1.0}


def reduce_sum(ops):
    return functools.reduce(operator.add, ops)
