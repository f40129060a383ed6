"""Minimal function objects used as prior means / right-hand sides on the hot path.

Mirrors ``linpde_gp.functions.Zero/Constant`` (src/linpde_gp/functions/_constant.py:12,52) and the
``pn.functions.Function`` calling convention (pn/functions/_function.py:71-112): ``f(x)`` maps
``batch_shape + input_shape`` to ``batch_shape + output_shape``.
"""
from __future__ import annotations

import numpy as np


def _as_shape(shape) -> tuple:
    if shape is None:
        return ()
    if np.ndim(shape) == 0:
        return (int(shape),)
    return tuple(int(s) for s in shape)


class Function:
    def __init__(self, input_shape=(), output_shape=()):
        self._input_shape = _as_shape(input_shape)
        self._output_shape = _as_shape(output_shape)

    @property
    def input_shape(self):
        return self._input_shape

    @property
    def input_ndim(self):
        return len(self._input_shape)

    @property
    def output_shape(self):
        return self._output_shape

    @property
    def output_ndim(self):
        return len(self._output_shape)

    def __call__(self, x):
        x = np.asarray(x, dtype=np.double)
        if x.shape[x.ndim - self.input_ndim :] != self._input_shape:
            raise ValueError(
                f"The shape of the input {x.shape} is not compatible with the specified `input_shape` "
                f"of the `Function` {self._input_shape}."
            )
        return self._evaluate(x)

    def _evaluate(self, x):  # pragma: no cover - abstract
        raise NotImplementedError

    # -- arithmetic (pn/functions/_algebra_fallbacks.py: ScaledFunction / SumFunction) ---------------------------
    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledFunction(self, scalar=other)
        return NotImplemented

    def __neg__(self):
        return -1.0 * self

    def __add__(self, other):
        if isinstance(other, Function):
            return SumFunction(self, other)
        return NotImplemented

    def __sub__(self, other):
        if isinstance(other, Function):
            return SumFunction(self, -other)
        return NotImplemented


class ScaledFunction(Function):
    def __init__(self, function: Function, scalar):
        super().__init__(function.input_shape, function.output_shape)
        self._function = function
        self._scalar = float(scalar)

    @property
    def function(self):
        return self._function

    @property
    def scalar(self):
        return self._scalar

    def _evaluate(self, x):
        return self._scalar * self._function(x)


class SumFunction(Function):
    def __init__(self, *summands: Function):
        first = summands[0]
        if not all(s.input_shape == first.input_shape and s.output_shape == first.output_shape for s in summands):
            raise ValueError("The functions must have the same input and output shapes")
        super().__init__(first.input_shape, first.output_shape)
        self._summands = tuple(summands)

    @property
    def summands(self):
        return self._summands

    def _evaluate(self, x):
        out = self._summands[0](x)
        for s in self._summands[1:]:
            out = out + s(x)
        return out


class StackedFunction(Function):
    """``x -> (f_0(x), ..., f_{n-1}(x))`` of scalar-output functions (src/linpde_gp/functions/_stacked.py): the prior
    mean of a multi-output process."""

    def __init__(self, *fns: Function):
        if not fns:
            raise ValueError("at least one function is needed")
        if not all(f.input_shape == fns[0].input_shape and f.output_shape == () for f in fns):
            raise ValueError("stacked functions must share the input shape and be scalar-valued")
        super().__init__(fns[0].input_shape, (len(fns),))
        self._fns = tuple(fns)

    @property
    def fns(self):
        return self._fns

    def _evaluate(self, x):
        return np.stack([f(x) for f in self._fns], axis=-1)


class Constant(Function):
    def __init__(self, input_shape, value):
        self._value = np.asarray(value, dtype=np.double)
        super().__init__(input_shape, self._value.shape)

    @property
    def value(self):
        return self._value

    def _evaluate(self, x):
        batch = x.shape[: x.ndim - self.input_ndim]
        return np.broadcast_to(self._value, batch + self.output_shape).copy()


class Zero(Constant):
    def __init__(self, input_shape, output_shape=()):
        super().__init__(input_shape, np.zeros(_as_shape(output_shape)))


class LambdaFunction(Function):
    """Host callable wrapped as a Function (prior means for plain evaluation observations)."""

    def __init__(self, fn, input_shape=(), output_shape=()):
        super().__init__(input_shape, output_shape)
        self._fn = fn

    def _evaluate(self, x):
        return np.asarray(self._fn(x), dtype=np.double)


# -- univariate polynomials (src/linpde_gp/functions/_polynomial.py:17-238) -----------------------------------------------
class Monomial(Function):
    """``x -> x^degree`` on scalar inputs."""

    def __init__(self, degree: int):
        super().__init__(input_shape=(), output_shape=())
        degree = int(degree)
        if degree < 0:
            raise ValueError("The degree of the monomial must be non-negative.")
        self._degree = degree

    @property
    def degree(self) -> int:
        return self._degree

    def _evaluate(self, x):
        return x**self._degree


class Polynomial(Function):
    """``x -> sum_i coeffs[i] x^i`` (ascending coefficients) on scalar inputs, evaluated by Horner's rule
    (_polynomial.py:39-96).  ``differentiate`` / ``integrate`` are exact on the coefficients; differential operators
    and ``LebesgueIntegral`` apply to polynomials in closed form (``LinearFunctionOperator._apply_to_function``,
    ``linfunctls._integrate_function``), so that polynomial prior means and right-hand sides stay on the closed-form
    path (the reference differentiates them through its JAX fallback)."""

    def __init__(self, coeffs):
        super().__init__(input_shape=(), output_shape=())
        coeffs = tuple(float(c) for c in coeffs)
        self._coeffs = coeffs if len(coeffs) >= 1 else (0.0,)

    @property
    def coefficients(self) -> tuple:
        return self._coeffs

    @property
    def degree(self) -> int:
        return len(self._coeffs) - 1

    def __repr__(self) -> str:
        return " + ".join(str(c) if k == 0 else f"{c} x^{k}" for k, c in enumerate(self._coeffs) if c != 0.0) or "0"

    def _evaluate(self, x):
        res = np.full_like(x, self._coeffs[-1], dtype=np.double)
        for c in self._coeffs[-2::-1]:
            res = res * x + c
        return res

    def _new(self, coeffs):
        return Polynomial(coeffs)

    def differentiate(self):
        return self._new(tuple(c * k for k, c in enumerate(self._coeffs[1:], start=1)))

    def integrate(self):
        return self._new((0 * self._coeffs[0],) + tuple(c / (i + 1) for i, c in enumerate(self._coeffs)))

    def __neg__(self):
        return self._new(tuple(-c for c in self._coeffs))

    def _combine(self, other, sign):
        import itertools  # pylint: disable=import-outside-toplevel

        both_rational = isinstance(self, RationalPolynomial) and isinstance(other, RationalPolynomial)
        if both_rational:  # exact
            a, b, zero = self._coeffs, other._coeffs, 0 * self._coeffs[0]  # pylint: disable=protected-access
        else:
            a, b, zero = self.coefficients, other.coefficients, 0.0
        coeffs = tuple(c0 + sign * c1 for c0, c1 in itertools.zip_longest(a, b, fillvalue=zero))
        return RationalPolynomial(coeffs) if both_rational else Polynomial(coeffs)

    def __add__(self, other):
        if isinstance(other, Polynomial):
            return self._combine(other, 1)
        if isinstance(other, Constant) and other.input_shape == () and other.output_shape == ():
            return Polynomial((float(self._coeffs[0]) + float(other.value),) + tuple(float(c) for c in self._coeffs[1:]))
        return super().__add__(other)

    def __sub__(self, other):
        if isinstance(other, Polynomial):
            return self._combine(other, -1)
        return super().__sub__(other)

    def __rmul__(self, other):
        if np.ndim(other) == 0 and not isinstance(other, Function):
            return Polynomial(tuple(float(other) * float(c) for c in self._coeffs))
        return super().__rmul__(other)

    def __floordiv__(self, other):
        if not isinstance(other, Monomial):
            return NotImplemented
        if not 0 <= other.degree <= self.degree:
            raise ValueError("The degree of the monomial is larger than the degree of the polynomial")
        if any(c != 0 for c in self._coeffs[: other.degree]):
            raise ValueError(f"The first {other.degree} of the polynomial are not all zeros")
        return self._new(self._coeffs[other.degree:])


class RationalPolynomial(Polynomial):
    """Polynomial with exact ``fractions.Fraction`` coefficients (_polynomial.py:166-238): the derivative polynomials
    of half-integer Matern kernels are kept exact until they are folded into the device descriptor."""

    def __init__(self, coeffs):
        from fractions import Fraction  # pylint: disable=import-outside-toplevel

        coeffs = tuple(Fraction(c) for c in coeffs)
        if len(coeffs) < 1:
            coeffs = (Fraction(0),)
        Function.__init__(self, input_shape=(), output_shape=())
        self._coeffs = coeffs

    @property
    def rational_coefficients(self) -> tuple:
        return self._coeffs

    @property
    def coefficients(self) -> tuple:
        return tuple(float(c) for c in self._coeffs)

    def _new(self, coeffs):
        return RationalPolynomial(coeffs)

    def _evaluate(self, x):
        cs = self.coefficients
        res = np.full_like(x, cs[-1], dtype=np.double)
        for c in cs[-2::-1]:
            res = res * x + c
        return res

    def __repr__(self) -> str:
        if all(c == 0 for c in self._coeffs):
            return "0"
        return " ".join(
            str(c) if k == 0 else "".join(["+" if c > 0 else "-", f" {abs(c)}" if abs(c) != 1 else "", f" x^{k}"])
            for k, c in enumerate(self._coeffs) if c != 0)

    def __rmul__(self, other):
        from fractions import Fraction  # pylint: disable=import-outside-toplevel

        if isinstance(other, (int, Fraction)) and not isinstance(other, bool):
            return RationalPolynomial(tuple(Fraction(other) * c for c in self._coeffs))
        return super().__rmul__(other)


class TruncatedSineSeries(Function):
    """``x -> sum_n c_n sin(n pi (x - l) / (r - l))`` on an interval ``[l, r]`` (src/linpde_gp/functions/_fourier.py:12-63):
    the initial values of the heat-equation problems."""

    def __init__(self, domain, coefficients):
        from .. import domains  # pylint: disable=import-outside-toplevel

        domain = domains.asdomain(domain)
        if not isinstance(domain, domains.Interval):
            raise TypeError("`domain` must be an `Interval`")
        self._domain = domain
        super().__init__(domain.shape, ())
        coefficients = np.asarray(coefficients, dtype=np.double)
        if coefficients.ndim != 1:
            raise ValueError("`coefficients` must be one-dimensional")
        self._coefficients = coefficients

    domain = property(lambda self: self._domain)
    coefficients = property(lambda self: self._coefficients)

    @property
    def half_angular_frequencies(self) -> np.ndarray:
        l, r = self._domain
        return np.pi * np.arange(1, self._coefficients.shape[-1] + 1) / (r - l)

    def _evaluate(self, x):
        l, _ = self._domain
        return np.sum(self._coefficients * np.sin(self.half_angular_frequencies * (x[..., None] - l)), axis=-1)


from . import bases  # noqa: E402,F401  (piecewise-linear finite-element basis; imports Function from this module)
