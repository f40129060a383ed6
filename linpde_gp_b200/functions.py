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
