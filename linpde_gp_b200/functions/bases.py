"""Piecewise-linear ("hat") finite-element basis on a 1-D grid: ``linpde_gp.functions.bases``.

``UnivariateLinearInterpolationBasis`` follows src/linpde_gp/functions/bases/_fem.py:7-117: with
``zero_boundary=False`` two sentinel nodes are added and the first / last basis functions keep only the half of their
support inside the original grid; ``l2_projection()`` returns the functional of
``linpde_gp_b200/linfunctls/projections/l2.py``.  Host set-up code (grids of tens to thousands of nodes)."""
from __future__ import annotations

import numpy as np

from . import Function


class UnivariateLinearInterpolationBasis(Function):
    def __init__(self, grid, zero_boundary: bool = False) -> None:
        grid = np.asarray(grid, dtype=np.double)
        zero_boundary = bool(zero_boundary)
        if grid.ndim != 1 or grid.size < 3:
            raise ValueError("`grid` must be a one-dimensional array of at least 3 nodes")
        if not np.all(np.diff(grid) > 0):
            raise ValueError("`grid` must be strictly increasing")
        if not zero_boundary:  # sentinel nodes (_fem.py:17-25)
            grid = np.concatenate(([grid[0] - (grid[1] - grid[0])], grid, [grid[-1] + (grid[-1] - grid[-2])]))
        self._grid = grid
        self._zero_boundary = zero_boundary
        self._left_normalization_factors = 1.0 / (self.x_i - self.x_im1)
        self._right_normalization_factors = 1.0 / (self.x_ip1 - self.x_i)
        super().__init__(input_shape=(), output_shape=(self._grid.size - 2,))

    grid = property(lambda self: self._grid)
    x_im1 = property(lambda self: self._grid[:-2])
    x_i = property(lambda self: self._grid[1:-1])
    x_ip1 = property(lambda self: self._grid[2:])
    zero_boundary = property(lambda self: self._zero_boundary)

    def _evaluate(self, x):
        x = np.asarray(x, dtype=np.double)
        res = np.maximum(
            0.0,
            np.where(
                x[..., None] < self.x_i,
                (x[..., None] - self.x_im1) * self._left_normalization_factors,
                (self.x_ip1 - x[..., None]) * self._right_normalization_factors,
            ),
        )
        if not self._zero_boundary:
            res[x < self._grid[1], 0] = 0.0
            res[x > self._grid[-2], -1] = 0.0
        return res

    def eval_elem(self, idx: int, x):
        x = np.asarray(x, dtype=np.double)
        res = np.asarray(np.maximum(
            0.0,
            np.where(
                x < self.x_i[idx],
                (x - self.x_im1[idx]) * self._left_normalization_factors[idx],
                (self.x_ip1[idx] - x) * self._right_normalization_factors[idx],
            ),
        ))
        if not self._zero_boundary:
            res[x < self._grid[1]] = 0.0
            res[x > self._grid[-2]] = 0.0
        return res

    def support_bounds(self, idx: int):
        if not -len(self) <= idx < len(self):
            raise IndexError(idx)
        if not self._zero_boundary:
            if idx in (0, -len(self)):
                return self.x_i[0], self.x_ip1[0]
            if idx in (len(self) - 1, -1):
                return self.x_im1[-1], self.x_i[-1]
        return self.x_im1[idx], self.x_ip1[idx]

    def __len__(self):
        return self._output_shape[0]

    def l2_projection(self, normalized: bool = True):
        from ..linfunctls.projections.l2 import L2Projection_UnivariateLinearInterpolationBasis  # pylint: disable=import-outside-toplevel

        return L2Projection_UnivariateLinearInterpolationBasis(self, normalized=normalized)

    # -- elements and Gauss-Legendre nodes (quadrature of smooth integrands element by element) -------------------
    def elements(self) -> np.ndarray:
        """Nodes bounding the elements that carry basis functions: the original grid (sentinels excluded)."""
        return self._grid if self._zero_boundary else self._grid[1:-1]

    def gauss_legendre(self, order: int = 20):
        """``(nodes, W)``: Gauss-Legendre nodes on every element and the ``len(self) x len(nodes)`` matrix
        ``W[i, q] = weight_q * phi_i(node_q)``, so that ``int phi_i f = W @ f(nodes)`` up to the quadrature error of a
        function that is smooth INSIDE each element."""
        t, w = np.polynomial.legendre.leggauss(int(order))
        e = self.elements()
        mid, half = 0.5 * (e[1:] + e[:-1]), 0.5 * (e[1:] - e[:-1])
        nodes = (mid[:, None] + half[:, None] * t[None, :]).reshape(-1)
        weights = (half[:, None] * w[None, :]).reshape(-1)
        return nodes, (self._evaluate(nodes) * weights[:, None]).T.copy()
