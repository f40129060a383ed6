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
