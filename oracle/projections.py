"""Oracle: L2 projections onto piecewise-linear bases (numpy/scipy restatement of the reference's algorithm).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows

* ``UnivariateLinearInterpolationBasis`` -- src/linpde_gp/functions/bases/_fem.py:7-117 (sentinel nodes :17-25,
  ``eval_elem`` :75-93, ``support_bounds`` :95-106);
* ``L2Projection_UnivariateLinearInterpolationBasis.normalizer`` -- src/linpde_gp/linfunctls/projections/l2/_fem.py:38-62;
* ``CovarianceFunction_L2Projection_UnivariateLinearInterpolationBasis._evaluate`` (scipy.integrate.quad per point and
  basis function) -- src/linpde_gp/randprocs/crosscov/linfunctls/projections.py:44-69;
* the covariance of two projections (scipy.integrate.dblquad per entry) -- projections.py:72-108;
* ``Matern32_L2Projection_UnivariateLinearInterpolationBasis._evaluate`` (closed form) -- projections.py:125-170.

Parity pinned by ``tests/golden/projections.npz`` (outputs of the real reference; ``oracle/make_golden.py``)."""
from __future__ import annotations

import numpy as np
import scipy.integrate

from . import covfuncs as ocf


class Basis:
    def __init__(self, grid, zero_boundary: bool = False):
        grid = np.asarray(grid, dtype=float)
        if not zero_boundary:
            grid = np.concatenate(([grid[0] - (grid[1] - grid[0])], grid, [grid[-1] + (grid[-1] - grid[-2])]))
        self.grid, self.zero_boundary = grid, bool(zero_boundary)
        self.x_im1, self.x_i, self.x_ip1 = grid[:-2], grid[1:-1], grid[2:]

    def __len__(self):
        return self.grid.size - 2

    def eval_elem(self, idx, x):
        x = np.asarray(x, dtype=float)
        res = np.asarray(np.maximum(0.0, np.where(x < self.x_i[idx], (x - self.x_im1[idx]) / (self.x_i[idx] - self.x_im1[idx]),
                                                  (self.x_ip1[idx] - x) / (self.x_ip1[idx] - self.x_i[idx]))))
        if not self.zero_boundary:
            res[x < self.grid[1]] = 0.0
            res[x > self.grid[-2]] = 0.0
        return res

    def support_bounds(self, idx):
        if not self.zero_boundary:
            if idx == 0:
                return self.x_i[0], self.x_ip1[0]
            if idx == len(self) - 1:
                return self.x_im1[-1], self.x_i[-1]
        return self.x_im1[idx], self.x_ip1[idx]


def mass_matrix(basis: Basis) -> np.ndarray:
    diag = (basis.x_ip1 - basis.x_im1) / 3.0
    off = (basis.x_ip1[:-1] - basis.x_i[:-1]) / 6.0
    if not basis.zero_boundary:
        diag[0] = (basis.x_ip1[0] - basis.x_i[0]) / 3.0
        diag[-1] = (basis.x_i[-1] - basis.x_im1[-1]) / 3.0
    return np.diag(diag) + np.diag(off, 1) + np.diag(off, -1)


def normalize(basis: Basis, res: np.ndarray, axis: int, normalized: bool = True) -> np.ndarray:
    if not normalized:
        return res
    return np.moveaxis(np.linalg.solve(mass_matrix(basis), np.moveaxis(res, axis, 0).reshape(len(basis), -1)).reshape(
        np.moveaxis(res, axis, 0).shape), 0, axis)


def _k(kernel):
    if "base" not in kernel:  # a bare base-kernel spec
        kernel = {"scale": None, "base": kernel}
    return lambda x, t: float(ocf.matrix(kernel, None, None, np.atleast_1d(float(x)), np.atleast_1d(float(t)))[0, 0])


def crosscov_quad(kernel, basis: Basis, xs, normalized: bool = True) -> np.ndarray:
    """(len(xs), m): int phi_j(t) k(x_i, t) dt by adaptive quadrature, then the normaliser (projections.py:44-69)."""
    k = _k(kernel)
    res = np.array([[scipy.integrate.quad(lambda t, j=j, x=x: float(basis.eval_elem(j, t)) * k(x, t), *basis.support_bounds(j))[0]
                     for j in range(len(basis))] for x in np.asarray(xs, dtype=float)])
    return normalize(basis, res, -1, normalized)


def crosscov_matern32(lengthscale: float, basis: Basis, xs, normalized: bool = True) -> np.ndarray:
    """Closed form for nu = 3/2 (projections.py:129-170)."""
    x = np.asarray(xs, dtype=float)[..., None]
    alpha = np.sqrt(3) / lengthscale

    def aux(a, b, t0, al):
        anti = lambda t: -((t - x + 2.0 / al) * (t - t0 + 1.0 / al) + 1 / al**2) * np.exp(-al * (t - x))  # noqa: E731
        return anti(b) - anti(a)

    xm, xi, xp = basis.x_im1, basis.x_i, basis.x_ip1
    left = (aux(np.maximum(xm, x), np.maximum(xi, x), xm, alpha) + aux(np.minimum(xm, x), np.minimum(xi, x), xm, -alpha)) / (xi - xm)
    right = -(aux(np.maximum(xi, x), np.maximum(xp, x), xp, alpha) + aux(np.minimum(xi, x), np.minimum(xp, x), xp, -alpha)) / (xp - xi)
    res = left + right
    if not basis.zero_boundary:
        res[..., 0] = right[..., 0]
        res[..., -1] = left[..., -1]
    return normalize(basis, res, -1, normalized)


def covariance_dblquad(kernel, basis0: Basis, basis1: Basis, normalized0: bool = True, normalized1: bool = True) -> np.ndarray:
    """(m0, m1): double integrals by scipy dblquad, normalised on both sides (projections.py:72-108)."""
    k = _k(kernel)
    res = np.array([[scipy.integrate.dblquad(lambda x1, x0, i=i, j=j: float(basis0.eval_elem(i, x0)) * k(x0, x1) * float(basis1.eval_elem(j, x1)),
                                             *basis0.support_bounds(i), *basis1.support_bounds(j))[0]
                     for j in range(len(basis1))] for i in range(len(basis0))])
    res = normalize(basis1, res, -1, normalized1)
    return normalize(basis0, res, 0, normalized0)
