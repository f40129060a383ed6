"""Oracle: Lebesgue-integral observations of univariate half-integer Matern processes (numpy restatement).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows

* ``LebesgueIntegral`` -- src/linpde_gp/linfunctls/_integrals.py:13-62 (integral of a ``Constant`` = value * volume);
* ``HalfIntegerMaternRadialAntiderivative`` / ``...SecondAntiderivative`` --
  src/linpde_gp/randprocs/crosscov/linfunctls/integrals/_matern_lebesgue.py:14-108;
* ``UnivariateRadialCovarianceFunctionLebesgueIntegral._evaluate`` and
  ``univariate_radial_covfunc_lebesgue_integral_lebesgue_integral`` --
  src/linpde_gp/randprocs/crosscov/linfunctls/integrals/_radial_lebesgue.py:37-69;
* dispatch -- src/linpde_gp/randprocs/covfuncs/linfunctls/_registry.py:157-193.

Parity pinned by ``tests/golden/integrals.npz`` (outputs of the real reference on its own test cases,
tests/linpde_gp/randprocs/{crosscov,cov}/linfunctls/cases/cases_integral_matern.py; ``oracle/make_golden.py``).
"""
from __future__ import annotations

import functools
from fractions import Fraction

import numpy as np

from . import covfuncs as ocf


def _deriv(poly):
    return tuple(c * k for k, c in enumerate(poly[1:], start=1))


def _padd(a, b):
    n = max(len(a), len(b))
    a = tuple(a) + (Fraction(0),) * (n - len(a))
    b = tuple(b) + (Fraction(0),) * (n - len(b))
    return tuple(x + y for x, y in zip(a, b))


@functools.lru_cache(maxsize=None)
def antiderivative_polynomial(p: int):
    """sum_{m=0}^{p} P^{(m)} with P the Matern polynomial (_matern_lebesgue.py:22-34), exact rationals."""
    p_i = ocf.matern_half_integer_coefficients(p)
    poly = p_i
    for _ in range(p):
        p_i = _deriv(p_i)
        poly = _padd(poly, p_i)
    return poly


@functools.lru_cache(maxsize=None)
def second_antiderivative_polynomial(p: int):
    """P + sum_{i=1}^{p} (i + 1) P^{(i)} (_matern_lebesgue.py:71-83)."""
    p_i = ocf.matern_half_integer_coefficients(p)
    poly = p_i
    for i in range(1, p + 1):
        p_i = _deriv(p_i)
        poly = _padd(poly, tuple((i + 1) * c for c in p_i))
    return poly


def radial_antiderivative(p: int, r):
    """F(r) = int_0^r kappa_p(t) dt for the unit-lengthscale radial Matern profile (_matern_lebesgue.py:38-46)."""
    r = np.asarray(r, dtype=np.double)
    s = np.sqrt(2 * p + 1)
    poly = antiderivative_polynomial(p)
    c1 = (1.0 / s) * float(poly[0])
    return -(1.0 / s) * np.exp(-s * r) * ocf.horner(poly, s * r) + c1


def radial_second_antiderivative(p: int, r):
    """G(r) = int_0^r F(t) dt (_matern_lebesgue.py:87-95)."""
    r = np.asarray(r, dtype=np.double)
    s = np.sqrt(2 * p + 1)
    inv_2nu = 1.0 / (2 * p + 1)
    c1 = (1.0 / s) * float(antiderivative_polynomial(p)[0])
    poly = second_antiderivative_polynomial(p)
    c2 = -inv_2nu * float(poly[0])
    return inv_2nu * np.exp(-s * r) * ocf.horner(poly, s * r) + c1 * r + c2


def matern_lebesgue_integral(p: int, lengthscale: float, a: float, b: float, x):
    """x -> int_a^b k(x, t) dt for the univariate Matern-(p + 1/2) kernel (_radial_lebesgue.py:37-45)."""
    x = np.asarray(x, dtype=np.double)
    l = float(lengthscale)
    return l * (
        (-1.0) ** (b < x) * radial_antiderivative(p, np.abs(b - x) / l)
        - (-1.0) ** (a < x) * radial_antiderivative(p, np.abs(a - x) / l)
    )


def matern_lebesgue_integral_lebesgue_integral(p: int, lengthscale: float, dom0, dom1) -> float:
    """int_a^b int_c^d k(s, t) dt ds (_radial_lebesgue.py:54-69)."""
    l = float(lengthscale)
    (a, b), (c, d) = dom0, dom1
    g = functools.partial(radial_second_antiderivative, p)
    return float(l**2 * (g(abs(b - c) / l) - g(abs(a - c) / l) - g(abs(b - d) / l) + g(abs(a - d) / l)))


def _scaled_matern(kernel):
    """(scale, p, lengthscale) of an oracle kernel spec ``{"scale": s, "base": matern}`` with scalar input."""
    base = kernel["base"]
    if base["kind"] != "matern" or tuple(base.get("input_shape", ())) != ():
        raise NotImplementedError("closed-form integrals exist for univariate half-integer Matern kernels only")
    p = base["nu"] - 0.5
    if p != int(p):
        raise NotImplementedError("half-integer Matern only")
    scale = 1.0 if kernel.get("scale") is None else float(kernel["scale"])
    return scale, int(p), float(base["lengthscales"])


def integral_crosscov(kernel, dom, x):
    """(k int^*)(x) = int_dom k(x, t) dt for a (scaled) univariate Matern kernel spec."""
    scale, p, l = _scaled_matern(kernel)
    return scale * matern_lebesgue_integral(p, l, float(dom[0]), float(dom[1]), x)


def integral_integral(kernel, dom0, dom1) -> float:
    scale, p, l = _scaled_matern(kernel)
    return scale * matern_lebesgue_integral_lebesgue_integral(p, l, dom0, dom1)
