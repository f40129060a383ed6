"""Oracle: covariance functions and their operator-transformed versions (numpy restatement).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every evaluator here reproduces the sequence of
numpy operations of the reference class it cites, so that results agree with the reference to the last
few ulps; ``tests/test_oracle_golden.py`` pins them against outputs of the real reference.

Kernels are described by plain dicts (no product classes are imported here):

    {"kind": "matern",  "input_shape": () | (d,), "nu": 2.5, "lengthscales": float | (d,)}
    {"kind": "expquad", "input_shape": () | (d,), "lengthscales": float | (d,)}
    {"kind": "tensor_product", "factors": [<1-D matern/expquad dict>, ...]}

wrapped as ``{"scale": sigma2, "base": <kernel>}`` by :func:`scaled`.

Linear differential operators are lists of summands ``[(scalar, op), ...]`` (reference:
``SumLinearFunctionOperator`` of ``ScaledLinearDifferentialOperator``) with ``op`` one of

    ("pd", {multi_index_tuple: coeff, ...})   generic ``LinearDifferentialOperator.coefficients[()]``
    ("wl", weights)                           ``WeightedLaplacian(weights)``   (Laplacian = ones)
    ("dd", direction)                         ``DirectionalDerivative(direction)``

``None`` stands for the identity (no operator on that argument).
"""
from __future__ import annotations

import functools
from fractions import Fraction

import numpy as np

# --------------------------------------------------------------------------------------
# operators
# --------------------------------------------------------------------------------------


def laplacian(d: int, scalar: float = 1.0):
    """``scalar * Laplacian((d,))``  (src/linpde_gp/linfuncops/diffops/_laplacian.py:77-79)."""
    return [(float(scalar), ("wl", np.ones((d,) if d else (), dtype=np.double)))]


def weighted_laplacian(weights, scalar: float = 1.0):
    return [(float(scalar), ("wl", np.asarray(weights, dtype=np.double)))]


def directional_derivative(direction, scalar: float = 1.0):
    return [(float(scalar), ("dd", np.asarray(direction, dtype=np.double)))]


def heat_operator(d: int, alpha: float):
    """``HeatOperator((d,), alpha)`` = TimeDerivative + WeightedLaplacian([0, -a, ...])
    (src/linpde_gp/linfuncops/diffops/_heat.py:14-31)."""
    w = np.zeros((d,), dtype=np.double)
    w[1:] = -float(alpha)
    td = tuple([1] + [0] * (d - 1))
    return [(1.0, ("pd", {td: 1.0})), (1.0, ("wl", w))]


def op_coefficients(op) -> dict:
    """``LinearDifferentialOperator.coefficients[()]`` as ``{multi_index_tuple: coeff}`` in the reference's
    insertion order (``_laplacian.py:31-41``, ``_directional_derivative.py:19-30``: zero entries dropped)."""
    kind, payload = op
    if kind == "pd":
        return dict(payload)
    arr = np.asarray(payload, dtype=np.double)
    order = 2 if kind == "wl" else 1
    out = {}
    for idx, c in np.ndenumerate(arr):
        if c != 0.0:
            mi = np.zeros(arr.shape, dtype=int)
            mi[idx] = order
            out[tuple(int(v) for v in np.atleast_1d(mi))] = float(c)
    return out


# --------------------------------------------------------------------------------------
# base kernels
# --------------------------------------------------------------------------------------


@functools.lru_cache(maxsize=None)
def matern_half_integer_coefficients(p: int):
    """pn/randprocs/covfuncs/_matern.py:227-262 (ascending powers, exact rationals)."""
    coeffs = [Fraction(1, 1)]
    for i in range(p - 1, -1, -1):
        coeffs.append(coeffs[-1] * 2 * (i + 1) / (p + i + 1) / (p - i))
    return tuple(coeffs)


@functools.lru_cache(maxsize=None)
def matern_derivative_polynomial(p: int, n: int):
    """``half_integer_matern_derivative_polynomial``: P_{p,n} = P'_{p,n-1} - P_{p,n-1}
    (src/linpde_gp/randprocs/covfuncs/linfuncops/diffops/_matern.py:613-639)."""
    if n == 0:
        return matern_half_integer_coefficients(p)
    prev = matern_derivative_polynomial(p, n - 1)
    deriv = tuple(c * k for k, c in enumerate(prev[1:], start=1))
    deriv = deriv + (Fraction(0),) * (len(prev) - len(deriv))
    return tuple(d - c for d, c in zip(deriv, prev))


def _poly_floordiv_monomial(coeffs, degree: int):
    """``Polynomial.__floordiv__(Monomial)`` (src/linpde_gp/functions/_polynomial.py:148-163)."""
    if any(c != 0 for c in coeffs[:degree]):
        raise ValueError(f"The first {degree} coefficients of the polynomial are not all zeros")
    return tuple(coeffs[degree:])


def horner(coeffs, x: np.ndarray) -> np.ndarray:
    """``Polynomial._evaluate`` (src/linpde_gp/functions/_polynomial.py:61-68)."""
    coeffs = tuple(float(c) for c in coeffs) or (0.0,)
    res = np.full_like(x, coeffs[-1])
    for k in range(len(coeffs) - 2, -1, -1):
        res *= x
        res += coeffs[k]
    return res


def _input_ndim(k) -> int:
    return len(tuple(k.get("input_shape", ())))


def _batch_shape(k, x):
    nd = _input_ndim(k)
    return x.shape[: x.ndim - nd]


def _bsum(k, a):
    nd = _input_ndim(k)
    return np.sum(a, axis=tuple(range(-nd, 0)))


def _matern_p(k):
    p = k["nu"] - 0.5
    if p != int(p):
        raise NotImplementedError("oracle covers half-integer Matern only")
    return int(p)


def _matern_scale_factors(k):
    """pn/randprocs/covfuncs/_matern.py:172-173."""
    return np.sqrt(2 * k["nu"]) / np.asarray(k["lengthscales"], dtype=np.double)


def _expquad_scale_factors(k):
    """pn/randprocs/covfuncs/_exponentiated_quadratic.py:86-87."""
    return np.sqrt(0.5) / np.asarray(k["lengthscales"], dtype=np.double)


def _sq_euclid(k, x0, x1, scale_factors):
    """``IsotropicMixin._squared_euclidean_distances`` (pn/randprocs/covfuncs/_covariance_function.py:757-781)."""
    if x1 is None:
        return np.zeros(_batch_shape(k, x0), dtype=x0.dtype)
    sqdiffs = x0 - x1
    sqdiffs = sqdiffs * scale_factors
    sqdiffs *= sqdiffs
    return _bsum(k, sqdiffs)


def _euclid(k, x0, x1, scale_factors):
    """``IsotropicMixin._euclidean_distances`` (…/_covariance_function.py:783-802)."""
    if x1 is None:
        return np.zeros(_batch_shape(k, x0), dtype=x0.dtype)
    return np.sqrt(_sq_euclid(k, x0, x1, scale_factors))


def matern_evaluate(k, x0, x1):
    """``Matern._evaluate`` half-integer branch (pn/randprocs/covfuncs/_matern.py:175-195)."""
    p = _matern_p(k)
    scaled_dists = _euclid(k, x0, x1, _matern_scale_factors(k))
    coeffs = np.asarray(matern_half_integer_coefficients(p), dtype=np.float64)
    res = np.full_like(scaled_dists, coeffs[p])
    for i in range(p - 1, -1, -1):
        res *= scaled_dists
        res += coeffs[i]
    res *= np.exp(-scaled_dists)
    return res


def expquad_evaluate(k, x0, x1):
    """``ExpQuad._evaluate`` (pn/randprocs/covfuncs/_exponentiated_quadratic.py:89-100)."""
    if x1 is None:
        return np.ones(_batch_shape(k, x0), dtype=x0.dtype)
    return np.exp(-_sq_euclid(k, x0, x1, _expquad_scale_factors(k)))


# --------------------------------------------------------------------------------------
# half-integer Matern x {DirectionalDerivative, WeightedLaplacian}
# (src/linpde_gp/randprocs/covfuncs/linfuncops/diffops/_matern.py)
# --------------------------------------------------------------------------------------


def matern_id_dd(k, direction, reverse, x0, x1):
    """``HalfIntegerMatern_Identity_DirectionalDerivative._evaluate`` (_matern.py:17-86)."""
    if x1 is None:
        return np.zeros(_batch_shape(k, x0), dtype=x0.dtype)
    s = _matern_scale_factors(k)
    poly = _poly_floordiv_monomial(matern_derivative_polynomial(_matern_p(k), 1), 1)
    scaled_direction = s * np.asarray(direction, dtype=np.double)
    if not reverse:
        scaled_direction = scaled_direction * -1
    scaled_diffs = (x0 - x1) * s
    proj = _bsum(k, scaled_direction * scaled_diffs)
    dists = np.sqrt(_bsum(k, scaled_diffs**2))
    res = horner(poly, dists)
    res *= np.exp(-dists)
    res *= proj
    return res


def matern_dd_dd(k, direction0, direction1, x0, x1):
    """Univariate (``_matern.py:267-318``) or multivariate (``:138-264``) DD/DD closed form."""
    p = _matern_p(k)
    s = _matern_scale_factors(k)
    d0 = np.asarray(direction0, dtype=np.double)
    d1 = np.asarray(direction1, dtype=np.double)
    if int(np.prod(k.get("input_shape", ()))) == 1:
        poly = tuple(-c for c in matern_derivative_polynomial(p, 2))
        prod = np.squeeze(d0 * d1 * s**2)[()]
        if x1 is None:
            return np.full(_batch_shape(k, x0), prod * float(poly[0]), dtype=x0.dtype)
        dists = _euclid(k, x0, x1, s)
        return prod * horner(poly, dists) * np.exp(-dists)
    # _neg_poly_deriv = -P1 // r ;  _poly_diff = (P2 + _neg_poly_deriv) // r^2   (_matern.py:160-167)
    npd = _poly_floordiv_monomial(tuple(-c for c in matern_derivative_polynomial(p, 1)), 1)
    p2 = matern_derivative_polynomial(p, 2)
    length = max(len(p2), len(npd))
    tot = tuple((p2[i] if i < len(p2) else 0) + (npd[i] if i < len(npd) else 0) for i in range(length))
    poly_diff = _poly_floordiv_monomial(tot, 2)
    sd0 = s * d0
    sd1 = s * d1
    inprod = np.sum(sd0 * sd1)
    if x1 is None:
        return np.full(_batch_shape(k, x0), inprod * float(npd[0]), dtype=x0.dtype)
    scaled_diffs = (x0 - x1) * s
    proj0 = _bsum(k, sd0 * scaled_diffs)
    proj1 = _bsum(k, sd1 * scaled_diffs)
    dists = np.sqrt(_bsum(k, scaled_diffs**2))
    res = inprod * horner(npd, dists)
    res -= proj0 * proj1 * horner(poly_diff, dists)
    return res * np.exp(-dists)


def matern_id_wl(k, weights, x0, x1):
    """``UnivariateHalfIntegerMatern_Identity_WeightedLaplacian._evaluate`` (_matern.py:358-410)."""
    s = _matern_scale_factors(k)
    poly = matern_derivative_polynomial(_matern_p(k), 2)
    osf = np.squeeze(np.asarray(weights, dtype=np.double) * s * s)[()]
    dists = _euclid(k, x0, x1, s)
    return osf * np.exp(-dists) * horner(poly, dists)


def matern_wl_wl(k, weights0, weights1, x0, x1):
    """``UnivariateHalfIntegerMatern_WeightedLaplacian_WeightedLaplacian._evaluate`` (_matern.py:439-483)."""
    s = _matern_scale_factors(k)
    poly = matern_derivative_polynomial(_matern_p(k), 4)
    osf = np.squeeze(np.asarray(weights0, dtype=np.double) * np.asarray(weights1, dtype=np.double) * s**4)[()]
    dists = _euclid(k, x0, x1, s)
    return osf * np.exp(-dists) * horner(poly, dists)


def matern_dd_wl(k, direction, weights, reverse, x0, x1):
    """``UnivariateHalfIntegerMatern_DirectionalDerivative_WeightedLaplacian._evaluate`` (_matern.py:512-571)."""
    if x1 is None:
        return np.zeros(_batch_shape(k, x0), dtype=x0.dtype)
    s = _matern_scale_factors(k)
    poly = _poly_floordiv_monomial(matern_derivative_polynomial(_matern_p(k), 3), 1)
    scaled_direction = np.asarray(weights, dtype=np.double) * s**3
    scaled_direction = scaled_direction * np.asarray(direction, dtype=np.double)
    if reverse:
        scaled_direction = scaled_direction * -1
    scaled_diffs = (x0 - x1) * s
    proj = _bsum(k, scaled_direction * scaled_diffs)
    dists = np.sqrt(_bsum(k, scaled_diffs**2))
    return np.exp(-dists) * horner(poly, dists) * proj


# --------------------------------------------------------------------------------------
# ExpQuad x {DirectionalDerivative, WeightedLaplacian}
# (src/linpde_gp/randprocs/covfuncs/linfuncops/diffops/_expquad.py)
# --------------------------------------------------------------------------------------


def expquad_id_dd(k, direction, reverse, x0, x1):
    """``ExpQuad_Identity_DirectionalDerivative._evaluate`` (_expquad.py:12-73)."""
    if x1 is None:
        return np.zeros(_batch_shape(k, x0), dtype=x0.dtype)
    ell = np.asarray(k["lengthscales"], dtype=np.double)
    rescaled = np.asarray(direction, dtype=np.double) / ell**2
    if reverse:
        rescaled = -rescaled
    diffs = x0 - x1
    proj = _bsum(k, rescaled * diffs)
    dists_sq = _bsum(k, (diffs / ell) ** 2)
    return proj * np.exp(-0.5 * dists_sq)


def expquad_dd_dd(k, direction0, direction1, x0, x1):
    """``ExpQuad_DirectionalDerivative_DirectionalDerivative._evaluate`` (_expquad.py:76-142)."""
    ell = np.asarray(k["lengthscales"], dtype=np.double)
    d0 = np.asarray(direction0, dtype=np.double)
    r0 = d0 / ell**2
    r1 = np.asarray(direction1, dtype=np.double) / ell**2
    inprod = _bsum(k, d0 * r1) if _input_ndim(k) else d0 * r1
    if x1 is None:
        return np.full(_batch_shape(k, x0), inprod, dtype=x0.dtype)
    diffs = x0 - x1
    proj0 = _bsum(k, r0 * diffs)
    proj1 = _bsum(k, r1 * diffs)
    dists_sq = _bsum(k, (diffs / ell) ** 2)
    return (inprod - proj0 * proj1) * np.exp(-0.5 * dists_sq)


def _eq_wl_terms(k, weights):
    sf = _expquad_scale_factors(k)
    w = np.asarray(weights, dtype=np.double)
    weighted_inv_ls_sq = 2.0 * w * sf**2  # _expquad.py:171-173
    scale_factors_sq = 2.0 * weighted_inv_ls_sq * sf**2  # :175-179
    trace_term = np.sum(2.0 * w * sf**2)  # :181-183
    return sf, weighted_inv_ls_sq, scale_factors_sq, trace_term


def expquad_id_wl(k, weights, x0, x1):
    """``ExpQuad_Identity_WeightedLaplacian._evaluate`` (_expquad.py:145-201)."""
    sf, _, sfsq, trace = _eq_wl_terms(k, weights)
    if x1 is None:
        return np.full(_batch_shape(k, x0), -trace, dtype=x0.dtype)
    diffs = x0 - x1
    return (_bsum(k, sfsq * diffs * diffs) - trace) * np.exp(-_bsum(k, (sf * diffs) ** 2))


def expquad_wl_wl(k, weights0, weights1, x0, x1):
    """``ExpQuad_WeightedLaplacian_WeightedLaplacian._evaluate`` (_expquad.py:221-312)."""
    sf, wils0, sfsq0, trace0 = _eq_wl_terms(k, weights0)
    _, wils1, sfsq1, trace1 = _eq_wl_terms(k, weights1)
    sfsq = sfsq0 * wils1  # :266-271
    trace = np.sum(wils0 * wils1)  # :273-278
    if x1 is None:
        return np.full(_batch_shape(k, x0), trace0 * trace1 + 2 * trace, dtype=x0.dtype)
    diffs = x0 - x1
    return (
        (_bsum(k, sfsq0 * diffs**2) - trace0) * (_bsum(k, sfsq1 * diffs**2) - trace1)
        - 4 * _bsum(k, sfsq * diffs**2)
        + 2 * trace
    ) * np.exp(-_bsum(k, (sf * diffs) ** 2))


def expquad_dd_wl(k, direction, weights, reverse, x0, x1):
    """``ExpQuad_DirectionalDerivative_WeightedLaplacian._evaluate`` (_expquad.py:348-432).

    ``reverse=False``: direction on argument 0, Laplacian on argument 1."""
    if x1 is None:
        return np.zeros(_batch_shape(k, x0), dtype=x0.dtype)
    sf, wils, _, _ = _eq_wl_terms(k, weights)
    rescaled_direction = np.asarray(direction, dtype=np.double) * 2.0 * sf**2
    rescaled_weighted_direction = rescaled_direction * wils
    diffs = x0 - x1
    proj = _bsum(k, rescaled_direction * diffs)
    proj_w = _bsum(k, rescaled_weighted_direction * diffs)
    res = 2 * proj_w * expquad_evaluate(k, x0, x1)
    res -= proj * expquad_id_wl(k, weights, x0, x1)
    if reverse:
        return -res
    return res


# --------------------------------------------------------------------------------------
# 1-D factor dispatch:  PD_a(PD_b(k, argnum=1), argnum=0) for a, b in {0, 1, 2}
# (src/linpde_gp/randprocs/covfuncs/linfuncops/diffops/_registry.py:80-134 -> :142-370)
# --------------------------------------------------------------------------------------


def univariate_factor(k, a: int, b: int, x0, x1):
    """Value of the 1-D factor ``∂^a_{x0} ∂^b_{x1} k`` exactly as the reference's dispatcher builds it:
    order 1 -> ``DirectionalDerivative(1.0)``, order 2 -> ``WeightedLaplacian(1.0)``; argnum=1 first."""
    one = np.asarray(1.0)
    fam = k["kind"]
    if a > 2 or b > 2:
        raise NotImplementedError("reference has no closed form for 1-D orders >= 3 (_registry.py:88-96)")
    if fam == "matern":
        if (a, b) == (0, 0):
            return matern_evaluate(k, x0, x1)
        if (a, b) == (0, 1):
            return matern_id_dd(k, one, False, x0, x1)
        if (a, b) == (1, 0):
            return matern_id_dd(k, one, True, x0, x1)
        if (a, b) == (1, 1):
            return matern_dd_dd(k, one, one, x0, x1)
        if (a, b) in ((0, 2), (2, 0)):
            return matern_id_wl(k, one, x0, x1)
        if (a, b) == (2, 2):
            return matern_wl_wl(k, one, one, x0, x1)
        if (a, b) == (1, 2):  # L1 = Laplacian applied first (argnum=1), then direction on argnum=0
            return matern_dd_wl(k, one, one, False, x0, x1)
        if (a, b) == (2, 1):
            return matern_dd_wl(k, one, one, True, x0, x1)
    if fam == "expquad":
        if (a, b) == (0, 0):
            return expquad_evaluate(k, x0, x1)
        if (a, b) == (0, 1):
            return expquad_id_dd(k, one, False, x0, x1)
        if (a, b) == (1, 0):
            return expquad_id_dd(k, one, True, x0, x1)
        if (a, b) == (1, 1):
            return expquad_dd_dd(k, one, one, x0, x1)
        if (a, b) in ((0, 2), (2, 0)):
            return expquad_id_wl(k, one, x0, x1)
        if (a, b) == (2, 2):
            return expquad_wl_wl(k, one, one, x0, x1)
        if (a, b) == (1, 2):
            return expquad_dd_wl(k, one, one, False, x0, x1)
        if (a, b) == (2, 1):
            return expquad_dd_wl(k, one, one, True, x0, x1)
    raise NotImplementedError((fam, a, b))


def tensor_product_evaluate(k, x0, x1):
    """``TensorProduct._evaluate`` (src/linpde_gp/randprocs/covfuncs/_tensor_product.py:44-48, 85-95)."""
    res = None
    for i, f in enumerate(k["factors"]):
        v = univariate_factor(f, 0, 0, x0[..., i], x1[..., i] if x1 is not None else None)
        res = v if res is None else res * v
    return res


def tensor_product_lindiffop(k, coeffs0: dict, coeffs1: dict, x0, x1):
    """``TensorProduct_LinDiffOp_LinDiffOp._compute_res/_evaluate``
    (src/linpde_gp/randprocs/covfuncs/linfuncops/diffops/_tensor_product.py:84-119):
    res = Σ_{α∈L0} Σ_{β∈L1} c_α c_β Π_d factor_d^{(α_d, β_d)}, each distinct factor evaluated once."""
    d = len(k["factors"])
    memo = [dict() for _ in range(d)]
    res = 0.0
    for mi0, c0 in coeffs0.items():
        for mi1, c1 in coeffs1.items():
            factors = []
            for i in range(d):
                key = (mi0[i], mi1[i])
                if key not in memo[i]:
                    memo[i][key] = univariate_factor(
                        k["factors"][i], key[0], key[1], x0[..., i], x1[..., i] if x1 is not None else None
                    )
                factors.append(memo[i][key])
            res = res + c0 * c1 * functools.reduce(lambda u, v: u * v, factors)
    return res


# --------------------------------------------------------------------------------------
# L0 k L1*  for a (scaled) prior kernel
# --------------------------------------------------------------------------------------


def scaled(base, scale=None):
    return {"scale": None if scale is None else float(scale), "base": base}


def _identity_coeffs(d):
    return {tuple([0] * d): 1.0}


def _apply_pair(base, op0, op1, x0, x1):
    """One summand pair: base kernel with ``op0`` on argument 0 and ``op1`` on argument 1 (either may be
    ``None``); mirrors which closed-form class the reference's registry selects."""
    kind = base["kind"]
    if kind == "tensor_product":
        d = len(base["factors"])
        c0 = _identity_coeffs(d) if op0 is None else op_coefficients(op0)
        c1 = _identity_coeffs(d) if op1 is None else op_coefficients(op1)
        if op0 is None and op1 is None:
            return tensor_product_evaluate(base, x0, x1)
        return tensor_product_lindiffop(base, c0, c1, x0, x1)
    if op0 is None and op1 is None:
        return matern_evaluate(base, x0, x1) if kind == "matern" else expquad_evaluate(base, x0, x1)
    k0 = None if op0 is None else op0[0]
    k1 = None if op1 is None else op1[0]
    if "pd" in (k0, k1):
        raise NotImplementedError("generic partial derivatives have closed forms only on TensorProduct kernels")
    univariate = int(np.prod(base.get("input_shape", ()))) == 1
    if kind == "expquad":
        if k0 is None:
            return expquad_id_wl(base, op1[1], x0, x1) if k1 == "wl" else expquad_id_dd(base, op1[1], False, x0, x1)
        if k1 is None:
            return expquad_id_wl(base, op0[1], x0, x1) if k0 == "wl" else expquad_id_dd(base, op0[1], True, x0, x1)
        if (k0, k1) == ("wl", "wl"):
            return expquad_wl_wl(base, op0[1], op1[1], x0, x1)
        if (k0, k1) == ("dd", "dd"):
            return expquad_dd_dd(base, op0[1], op1[1], x0, x1)
        if (k0, k1) == ("dd", "wl"):
            return expquad_dd_wl(base, op0[1], op1[1], False, x0, x1)
        return expquad_dd_wl(base, op1[1], op0[1], True, x0, x1)
    if kind == "matern":
        if "wl" in (k0, k1) and not univariate:
            raise NotImplementedError("isotropic multi-d Matern x Laplacian is not closed-form in the reference (_registry.py:270-280)")
        if k0 is None:
            return matern_id_wl(base, op1[1], x0, x1) if k1 == "wl" else matern_id_dd(base, op1[1], False, x0, x1)
        if k1 is None:
            return matern_id_wl(base, op0[1], x0, x1) if k0 == "wl" else matern_id_dd(base, op0[1], True, x0, x1)
        if (k0, k1) == ("wl", "wl"):
            return matern_wl_wl(base, op0[1], op1[1], x0, x1)
        if (k0, k1) == ("dd", "dd"):
            return matern_dd_dd(base, op0[1], op1[1], x0, x1)
        if (k0, k1) == ("dd", "wl"):
            return matern_dd_wl(base, op0[1], op1[1], False, x0, x1)
        return matern_dd_wl(base, op1[1], op0[1], True, x0, x1)
    raise NotImplementedError(kind)


def evaluate(kernel, L0, L1, x0, x1):
    """Broadcast evaluation of ``L0 k L1*`` (``x1=None`` -> element-wise diagonal k(x0_i, x0_i)).

    Summation order follows the reference: ``L1`` is applied first (``argnum=1``,
    src/linpde_gp/randprocs/_gaussian_process/_lintransforms.py:12-13), then ``L0`` (``argnum=0``);
    sums are ``functools.reduce(operator.add)`` over summands
    (src/linpde_gp/linfuncops/_arithmetic.py:107-111; pn/randprocs/covfuncs/_arithmetic_fallbacks.py:67-112)
    and every scalar multiplies the already evaluated array
    (src/linpde_gp/randprocs/covfuncs/linfuncops/_registry.py:14-20)."""
    x0 = np.asarray(x0, dtype=np.double)
    x1 = None if x1 is None else np.asarray(x1, dtype=np.double)
    scale, base = kernel.get("scale"), kernel["base"]
    s0 = [(1.0, None)] if L0 is None else L0
    s1 = [(1.0, None)] if L1 is None else L1
    outer = None
    for c0, op0 in s0:
        inner = None
        for c1, op1 in s1:
            v = _apply_pair(base, op0, op1, x0, x1)
            if op1 is not None:
                v = c1 * v
            inner = v if inner is None else inner + v
        if op0 is not None:
            inner = c0 * inner
        outer = inner if outer is None else outer + inner
    if scale is not None:
        outer = scale * outer
    return outer


def matrix(kernel, L0, L1, x0, x1=None):
    """``CovarianceFunction._evaluate_matrix`` for scalar-output kernels
    (pn/randprocs/covfuncs/_covariance_function.py:553-582): K[i, j] = k(x0_i, x1_j); ``x1=None`` -> x1 := x0.
    Inputs are flattened C-order like ``_preprocess_linop_input`` (:676-693)."""
    base = kernel["base"]
    in_shape = (len(base["factors"]),) if base["kind"] == "tensor_product" else tuple(base.get("input_shape", ()))
    x0 = np.asarray(x0, dtype=np.double).reshape((-1,) + in_shape)
    x1 = x0 if x1 is None else np.asarray(x1, dtype=np.double).reshape((-1,) + in_shape)
    return evaluate(kernel, L0, L1, x0[(slice(None), None) + (Ellipsis,)], x1[(None, slice(None)) + (Ellipsis,)])


def matrix_tiled(kernel, L0, L1, x0, x1=None, tile_rows: int = 2048, out=None):
    """Row-tiled :func:`matrix` (the un-tiled broadcast needs ~10 N0xN1 temporaries)."""
    base = kernel["base"]
    in_shape = (len(base["factors"]),) if base["kind"] == "tensor_product" else tuple(base.get("input_shape", ()))
    x0 = np.asarray(x0, dtype=np.double).reshape((-1,) + in_shape)
    x1 = x0 if x1 is None else np.asarray(x1, dtype=np.double).reshape((-1,) + in_shape)
    if out is None:
        out = np.empty((x0.shape[0], x1.shape[0]), dtype=np.double)
    for r in range(0, x0.shape[0], tile_rows):
        out[r : r + tile_rows] = matrix(kernel, L0, L1, x0[r : r + tile_rows], x1)
    return out


def diagonal(kernel, L0, L1, x0):
    """k(x0_i, x0_i) via the reference's ``x1=None`` closed-form constants."""
    base = kernel["base"]
    in_shape = (len(base["factors"]),) if base["kind"] == "tensor_product" else tuple(base.get("input_shape", ()))
    x0 = np.asarray(x0, dtype=np.double).reshape((-1,) + in_shape)
    return np.broadcast_to(evaluate(kernel, L0, L1, x0, None), (x0.shape[0],)).copy()
