"""``LinearFunctionOperator`` and its arithmetic (src/linpde_gp/linfuncops/_linfuncop.py:16-105,
_arithmetic.py:12-160).  Operators are symbolic; calling one dispatches on the argument type."""
from __future__ import annotations

import functools
import operator

import numpy as np

from ..functions import Constant, Function, Polynomial, ScaledFunction, StackedFunction, SumFunction, Zero, _as_shape


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
        from ..randprocs import _conditional, _gaussian_process, covfuncs, crosscov

        if isinstance(f, covfuncs.CovarianceFunction):
            return covfuncs.apply_linfuncop(self, f, argnum=kwargs.get("argnum", 0))
        if isinstance(f, crosscov.ProcessVectorCrossCovariance):  # acts on the free argument (crosscov/linfuncops.py:18-87)
            return f._apply_linfuncop(self)  # pylint: disable=protected-access
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
        if isinstance(f, ScaledFunction):  # linearity (src/linpde_gp/functions/linfuncops/_registry.py)
            return f.scalar * self(f.function)
        if isinstance(f, SumFunction):
            return functools.reduce(operator.add, (self(s) for s in f.summands))
        terms = self._terms()
        order0 = sum(c for mi, c in terms.items() if sum(mi) == 0)
        if isinstance(f, Zero):
            return Zero(self.output_domain_shape, self.output_codomain_shape)
        if isinstance(f, Constant):
            if order0 == 0:  # pure derivative operator: Zero, as linfuncops/diffops/_functions.py:8-19
                return Zero(self.output_domain_shape, self.output_codomain_shape)
            return Constant(self.output_domain_shape, order0 * f.value)
        if isinstance(f, Polynomial) and self.input_domain_shape == () and self.input_codomain_shape == ():
            # univariate polynomial: sum_n c_n d^n/dx^n on the coefficients (exact for RationalPolynomial)
            res = None
            for mi, c in terms.items():
                g = f
                for _ in range(sum(mi)):
                    g = g.differentiate()
                g = Polynomial(tuple(float(c) * cc for cc in g.coefficients))
                res = g if res is None else res + g
            return res if res is not None else Zero((), ())
        raise NotImplementedError("only Zero / Constant / univariate Polynomial functions can be differentiated in "
                                  "closed form")

    def to_linfunctl(self, X):
        """``_EvaluationFunctional(X) @ self`` (``_linfuncop.py:93-105``)."""
        from .. import linfunctls

        return linfunctls.CompositeLinearFunctional(
            linop=None,
            linfunctl=linfunctls._EvaluationFunctional(  # pylint: disable=protected-access
                input_domain_shape=self.output_domain_shape,
                input_codomain_shape=self.output_codomain_shape,
                X=X,
            ),
            linfuncop=self,
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

    def __matmul__(self, other):
        """``self @ other``: apply ``other`` first (``_linfuncop.py:131-136``, ``_arithmetic.py:142-144``)."""
        if isinstance(other, SumLinearFunctionOperator):
            return SumLinearFunctionOperator(*(self @ summand for summand in other.summands))
        if isinstance(other, LinearFunctionOperator):
            return CompositeLinearFunctionOperator(self, other)
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

    def _apply_to_function(self, f):
        try:
            return super()._apply_to_function(f)
        except NotImplementedError:  # no flat partial-derivative form (e.g. a SelectOutput inside): linearity
            return float(self._scalar) * self._linfuncop(f)

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

    def _apply_to_function(self, f):
        try:
            return super()._apply_to_function(f)
        except NotImplementedError:
            return functools.reduce(operator.add, (s(f) for s in self._summands))

    def __repr__(self):
        return " + ".join(str(s) for s in self._summands)


class CompositeLinearFunctionOperator(LinearFunctionOperator):
    """``L_0 @ L_1 @ ...`` -- the rightmost operator acts first (src/linpde_gp/linfuncops/_arithmetic.py:112-139)."""

    def __init__(self, *linfuncops):
        assert all(L0.input_shapes == L1.output_shapes for L0, L1 in zip(linfuncops[:-1], linfuncops[1:]))
        self._linfuncops = tuple(linfuncops)
        super().__init__(input_shapes=self._linfuncops[-1].input_shapes, output_shapes=self._linfuncops[0].output_shapes)

    @property
    def linfuncops(self):
        return self._linfuncops

    def _terms(self):
        if len(self._linfuncops) == 1:
            return self._linfuncops[0]._terms()
        raise NotImplementedError("a composition has no flat partial-derivative representation")

    def _apply_to_function(self, f):
        return functools.reduce(lambda h, L: L(h), reversed(self._linfuncops), f)

    def __repr__(self):
        return " @ ".join(repr(L) for L in self._linfuncops)


class SelectOutput(LinearFunctionOperator):
    """Picks one output of a multi-output function / process (src/linpde_gp/linfuncops/_select_output.py:9-34)."""

    def __init__(self, input_shapes, idx):
        self._idx = idx
        in_dom, in_codom = _as_shape(input_shapes[0]), _as_shape(input_shapes[1])
        out_codom = np.empty(in_codom, dtype=[])[idx].shape
        super().__init__((in_dom, in_codom), output_shapes=(in_dom, out_codom))

    @property
    def idx(self):
        return self._idx

    def _terms(self):
        raise NotImplementedError("SelectOutput is not a differential operator")

    def _apply_to_function(self, f):
        if isinstance(f, StackedFunction):  # functions/linfuncops/_registry.py:8-13
            assert isinstance(self._idx, (int, np.integer))
            return f.fns[self._idx]
        if isinstance(f, ScaledFunction):
            return f.scalar * self(f.function)
        if isinstance(f, SumFunction):
            return functools.reduce(operator.add, (self(s) for s in f.summands))
        from ..functions import LambdaFunction

        idx = self._idx
        return LambdaFunction(lambda x: f(x)[..., idx], f.input_shape, self.output_codomain_shape)

    def __repr__(self):
        return f"SelectOutput(idx={self.idx})"


class Identity(LinearFunctionOperator):
    """Order-zero operator (the reference uses ``PartialDerivative(MultiIndex(zeros))``, diffops/_registry.py:41-50)."""

    def __init__(self, domain_shape, codomain_shape=()):
        super().__init__(input_shapes=(domain_shape, codomain_shape), output_shapes=(domain_shape, codomain_shape))

    def _terms(self):
        d = int(np.prod(self.input_domain_shape)) if self.input_domain_shape else 1
        return {tuple([0] * d): 1.0}


def reduce_sum(ops):
    return functools.reduce(operator.add, ops)
