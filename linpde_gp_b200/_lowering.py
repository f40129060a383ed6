"""Lowering of (prior kernel, L0, L1) to the flat ``lpgp_kernel_desc`` consumed by the CUDA kernels.

For stationary product kernels  k(x, x') = prod_d kappa_d(x_d - x'_d)  and operators given as sums of partial
derivatives  L = sum_alpha c_alpha d^alpha,

    (L0 k L1^*)(x, x') = sigma^2 sum_{alpha, beta} c_alpha c_beta prod_d (-1)^{beta_d} kappa_d^{(alpha_d + beta_d)}(delta_d)

(reference: src/linpde_gp/randprocs/covfuncs/linfuncops/diffops/_tensor_product.py:34-119).  Every 1-D factor is
"polynomial x exponential":

  Matern nu = p + 1/2 : kappa^{(n)}(delta) = s^n sign(delta)^n P_{p,n}(r) e^{-r},  r = s |delta|, s = sqrt(2 nu)/ell,
                        P_{p,n} = P'_{p,n-1} - P_{p,n-1}  (diffops/_matern.py:613-639)
  ExpQuad            : kappa^{(n)}(delta) = (-1)^n ell^{-n} He_n(v) e^{-v^2/2},   v = delta/ell  (diffops/_expquad.py)

so the whole double sum folds into ONE coefficient tensor over per-dimension monomial bases, evaluated on the
device with a single exp() per matrix entry.  The fold happens here, on the host, in exact rational arithmetic for
the polynomial parts.
"""
from __future__ import annotations

import functools
import itertools
from fractions import Fraction

import numpy as np

from . import _lib


@functools.lru_cache(maxsize=None)
def matern_poly(p: int, n: int):
    """Exact coefficients (ascending) of P_{p,n}: kappa_p^{(n)}(r) = P_{p,n}(r) e^{-r}."""
    if n == 0:
        c = [Fraction(1)]
        for i in range(p - 1, -1, -1):
            c.append(c[-1] * 2 * (i + 1) / (p + i + 1) / (p - i))
        return tuple(c)
    prev = matern_poly(p, n - 1)
    d = [prev[k] * k for k in range(1, len(prev))] + [Fraction(0)]
    return tuple(a - b for a, b in zip(d, prev))


@functools.lru_cache(maxsize=None)
def hermite_poly(n: int):
    """Probabilists' Hermite polynomial He_n (ascending coefficients): He_{n+1} = x He_n - n He_{n-1}."""
    if n == 0:
        return (Fraction(1),)
    if n == 1:
        return (Fraction(0), Fraction(1))
    a, b = hermite_poly(n - 1), hermite_poly(n - 2)
    out = [Fraction(0)] * (n + 1)
    for i, c in enumerate(a):
        out[i + 1] += c
    for i, c in enumerate(b):
        out[i] -= (n - 1) * c
    return tuple(out)


def _poly_deriv(c):
    return tuple(c[k] * k for k in range(1, len(c)))


def _poly_add(a, b):
    n = max(len(a), len(b))
    a = tuple(a) + (Fraction(0),) * (n - len(a))
    b = tuple(b) + (Fraction(0),) * (n - len(b))
    return tuple(x + y for x, y in zip(a, b))


@functools.lru_cache(maxsize=None)
def matern_antiderivative_polys(p: int):
    """Exact polynomial parts of the first / second radial antiderivative of a Matern-(p + 1/2) profile:
    P1 = sum_{m=0}^{p} P^(m), P2 = P + sum_{i=1}^{p} (i + 1) P^(i)
    (src/linpde_gp/randprocs/crosscov/linfunctls/integrals/_matern_lebesgue.py:22-34, 71-83)."""
    d = matern_poly(p, 0)
    p1 = p2 = d
    for i in range(1, p + 1):
        d = _poly_deriv(d)
        p1 = _poly_add(p1, d)
        p2 = _poly_add(p2, tuple((i + 1) * c for c in d))
    return p1, p2


def matern_integral_desc(nu: float, lengthscale: float) -> _lib.MaternIntegralDesc:
    """Descriptor of ``lpgp_matern_integral`` / ``lpgp_matern_integral2`` for a univariate half-integer Matern kernel."""
    p = nu - 0.5
    if p != int(p) or p < 0:
        raise NotImplementedError("closed-form Lebesgue integrals exist for half-integer Matern kernels only")
    p = int(p)
    if p + 1 > _lib.MAX_INTEGRAL_COEF:
        raise NotImplementedError(f"Matern order nu={nu} exceeds the device descriptor")
    p1, p2 = matern_antiderivative_polys(p)
    desc = _lib.MaternIntegralDesc()
    desc.ncoef = p + 1
    desc.scale = float(np.sqrt(2 * nu) / float(lengthscale))
    for i in range(p + 1):
        desc.poly1[i] = float(p1[i])
        desc.poly2[i] = float(p2[i])
    return desc


class Factor1D:
    """One stationary 1-D factor: ('matern', p, ell) or ('expquad', ell)."""

    def __init__(self, kind: str, lengthscale: float, nu: float | None = None):
        self.kind = kind
        self.lengthscale = float(lengthscale)
        self.nu = nu
        if kind == "matern":
            p = nu - 0.5
            if p != int(p) or p < 0:
                raise NotImplementedError("only half-integer Matern kernels have closed-form derivatives")
            self.p = int(p)
            self.scale = float(np.sqrt(2 * nu) / self.lengthscale)
        elif kind == "expquad":
            self.scale = 1.0 / self.lengthscale
        else:
            raise ValueError(kind)

    def deriv_coeffs(self, n: int):
        """(even, odd) coefficient vectors of kappa^{(n)} over [v^e] and [u v^e]."""
        if self.kind == "matern":
            P = matern_poly(self.p, n)
            sn = self.scale**n
            if n % 2 == 0:
                return [sn * float(c) for c in P], None
            if P[0] != 0:
                raise NotImplementedError(
                    f"derivative order {n} of a Matern-{self.nu} kernel is singular on the diagonal "
                    "(the reference raises for this combination as well)"
                )
            return None, [sn * float(c) for c in P[1:]] + [0.0]
        He = hermite_poly(n)
        f = (-1.0) ** n * self.scale**n
        return [f * float(c) for c in He], None


def lower(factors, L0, L1, sigma2: float = 1.0) -> _lib.KernelDesc:
    """Build the descriptor. ``L0``/``L1``: ``None`` (identity) or ``{multi_index_tuple: coeff}``."""
    d = len(factors)
    if not 1 <= d <= _lib.MAX_DIM:
        raise NotImplementedError(f"input dimension {d} not supported (max {_lib.MAX_DIM})")
    ident = {tuple([0] * d): 1.0}
    L0 = ident if L0 is None else L0
    L1 = ident if L1 is None else L1
    orders = [set() for _ in range(d)]
    for a, b in itertools.product(L0, L1):
        for i in range(d):
            orders[i].add(a[i] + b[i])
    nbasis, has_odd = [], []
    for i, f in enumerate(factors):
        if f.kind == "matern":
            nbasis.append(f.p + 1)
            has_odd.append(any(n % 2 for n in orders[i]))
        else:
            nbasis.append(max(orders[i]) + 1)
            has_odd.append(False)
    nbt = [nb * (2 if od else 1) for nb, od in zip(nbasis, has_odd)]
    if int(np.prod(nbt)) > _lib.MAX_COEF:
        raise NotImplementedError("coefficient tensor too large for the device descriptor")
    tensor = np.zeros(nbt, dtype=np.float64)
    cache = [{} for _ in range(d)]

    def vec(i, n):
        if n not in cache[i]:
            ev, od = factors[i].deriv_coeffs(n)
            v = np.zeros(nbt[i])
            if ev is not None:
                v[: len(ev)] = ev[: nbasis[i]] if len(ev) > nbasis[i] else ev
            if od is not None:
                v[nbasis[i] : nbasis[i] + len(od)] = od
            cache[i][n] = v
        return cache[i][n]

    for (a, ca), (b, cb) in itertools.product(L0.items(), L1.items()):
        term = np.asarray(float(ca) * float(cb))
        for i in range(d):
            sign = -1.0 if (b[i] % 2) else 1.0
            term = np.multiply.outer(term, sign * vec(i, a[i] + b[i]))
        tensor += term
    tensor *= float(sigma2)

    desc = _lib.KernelDesc()
    desc.d = d
    for i, f in enumerate(factors):
        desc.dim_type[i] = _lib.DIM_MATERN if f.kind == "matern" else _lib.DIM_EXPQUAD
        desc.nbasis[i] = nbasis[i]
        desc.has_odd[i] = int(has_odd[i])
        desc.scale[i] = f.scale
    flat = tensor.reshape(-1)
    for i, c in enumerate(flat):
        desc.coef[i] = float(c)
    desc.diag_value = float(flat[0])
    return desc


# ---------------------------------------------------------------------------------------------------------------
# radial kernels: isotropic multi-dimensional half-integer Matern and its first-order directional derivatives
# ---------------------------------------------------------------------------------------------------------------
def _floordiv_monomial(c, k: int):
    """Quotient of the polynomial ``c`` (ascending) by ``r^k`` (the reference's ``poly // Monomial(k)``)."""
    return tuple(c[k:]) if len(c) > k else (Fraction(0),)


def lower_radial(nu: float, scales, dir0=None, dir1=None, sigma2: float = 1.0) -> _lib.KernelDesc:
    """Descriptor of ``sigma2 * (D_{dir0} k D_{dir1}^*)(x, x')`` for the isotropic Matern-``nu`` kernel on R^d with
    per-dimension scale factors ``scales = sqrt(2 nu) / lengthscales``; ``dir0`` / ``dir1``: direction of the
    directional derivative on argument 0 / 1, or ``None`` (identity).  With u = s o (x - x'), r = |u|:

      identity / identity : P_p(r) e^{-r}                              (pn/randprocs/covfuncs/_matern.py:175-195)
      one derivative      : (P_{p,1} // r)(r) e^{-r} <+-s o dir, u>    (diffops/_matern.py:39-41, 55-86; minus sign when
                            the operator acts on argument 1)
      both                : [<a,b> N(r) - <a,u><b,u> Q(r)] e^{-r},  a = s o dir0, b = s o dir1,
                            N = -P_{p,1} // r,  Q = (P_{p,2} + N) // r^2                 (diffops/_matern.py:160-203)
    """
    p = nu - 0.5
    if p != int(p) or p < 0:
        raise NotImplementedError("only half-integer Matern kernels have closed-form derivatives")
    p = int(p)
    s = np.asarray(scales, dtype=np.float64).reshape(-1)
    d = s.size
    if not 1 <= d <= _lib.MAX_DIM:
        raise NotImplementedError(f"input dimension {d} not supported (max {_lib.MAX_DIM})")
    nq = _lib.RADIAL_NQ
    zero = (Fraction(0),)
    a = np.zeros(d)
    b = np.zeros(d)
    if dir0 is None and dir1 is None:
        q0, q1, q2, c0 = matern_poly(p, 0), zero, zero, 1.0
    elif dir0 is None or dir1 is None:
        if p < 1:
            raise NotImplementedError("the Matern-1/2 kernel is not differentiable")
        q0, q2, c0 = zero, zero, 1.0
        q1 = _floordiv_monomial(matern_poly(p, 1), 1)
        a = s * np.asarray(dir0, dtype=np.float64).reshape(-1) if dir1 is None else -s * np.asarray(dir1, dtype=np.float64).reshape(-1)
    else:
        if p < 1:
            raise NotImplementedError("the Matern-1/2 kernel is not differentiable")
        a = s * np.asarray(dir0, dtype=np.float64).reshape(-1)
        b = s * np.asarray(dir1, dtype=np.float64).reshape(-1)
        npd = _floordiv_monomial(tuple(-c for c in matern_poly(p, 1)), 1)
        q0, c0 = npd, float(np.sum(a * b))
        q1 = zero
        q2 = tuple(-c for c in _floordiv_monomial(_poly_add(matern_poly(p, 2), npd), 2))
    if a.size != d or b.size != d:
        raise ValueError("direction and kernel input dimensions differ")
    if max(len(q0), len(q1), len(q2)) > nq:
        raise NotImplementedError(f"Matern order nu={nu} exceeds the device descriptor")
    desc = _lib.KernelDesc()
    desc.d = d
    for i in range(d):
        desc.dim_type[i] = _lib.DIM_RADIAL
        desc.nbasis[i] = nq
        desc.has_odd[i] = 0
        desc.scale[i] = float(s[i])
    sig = float(sigma2)
    for i, c in enumerate(q0):
        desc.coef[i] = sig * c0 * float(c)
    for i, c in enumerate(q1):
        desc.coef[nq + i] = sig * float(c)
    for i, c in enumerate(q2):
        desc.coef[2 * nq + i] = sig * float(c)
    for i in range(d):
        desc.coef[3 * nq + i] = float(a[i])
        desc.coef[3 * nq + d + i] = float(b[i])
    desc.diag_value = sig * c0 * float(q0[0])
    return desc
