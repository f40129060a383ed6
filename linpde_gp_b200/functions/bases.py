"""Piecewise-linear ("hat") finite-element basis on a 1-D grid: ``linpde_gp.functions.bases``.

``UnivariateLinearInterpolationBasis`` follows src/linpde_gp/functions/bases/_fem.py:7-117: with
``zero_boundary=False`` two sentinel nodes are added and the first / last basis functions keep only the half of their
support inside the original grid; ``l2_projection()`` returns the functional of
``linpde_gp_b200/linfunctls/projections/l2.py``.  Host set-up code (grids of tens to thousands of nodes)."""
from __future__ import annotations

import numpy as np

from . import Function


class UnivariateLinearInterpolationBasis(Function):
    """``phi_i(x) = max(0, min(rise_i(x), fall_i(x)))`` with the two ramps through the neighbours of node ``c_i``.

    ``zero_boundary=True``: one function per INTERIOR node.  ``zero_boundary=False``: one function per node; the ramps of
    the first / last function are continued to a mirrored sentinel node and the function is cut off outside the grid, so
    that the basis is a partition of unity on ``[grid[0], grid[-1]]`` (the reference adds the same sentinels, which is why
    ``grid`` / ``x_im1`` / ``x_i`` / ``x_ip1`` below include them)."""

    def __init__(self, grid, zero_boundary: bool = False) -> None:
        nodes = np.array(grid, dtype=np.double)
        if nodes.ndim != 1 or nodes.size < 3:
            raise ValueError("`grid` must be a one-dimensional array of at least 3 nodes")
        steps = np.diff(nodes)
        if not np.all(steps > 0):
            raise ValueError("`grid` must be strictly increasing")
        self._zero_boundary = bool(zero_boundary)
        self._domain = (float(nodes[0]), float(nodes[-1]))
        if self._zero_boundary:
            self._grid = nodes
        else:
            self._grid = np.empty(nodes.size + 2)
            self._grid[1:-1] = nodes
            self._grid[0], self._grid[-1] = nodes[0] - steps[0], nodes[-1] + steps[-1]
        self._inv_rise = 1.0 / (self._grid[1:-1] - self._grid[:-2])
        self._inv_fall = 1.0 / (self._grid[2:] - self._grid[1:-1])
        super().__init__(input_shape=(), output_shape=(self._grid.size - 2,))

    grid = property(lambda self: self._grid)
    x_im1 = property(lambda self: self._grid[:-2])
    x_i = property(lambda self: self._grid[1:-1])
    x_ip1 = property(lambda self: self._grid[2:])
    zero_boundary = property(lambda self: self._zero_boundary)

    def _hats(self, x, which=slice(None)):
        x = np.asarray(x, dtype=np.double)[..., None]
        rise = (x - self.x_im1[which]) * self._inv_rise[which]
        fall = (self.x_ip1[which] - x) * self._inv_fall[which]
        vals = np.clip(np.minimum(rise, fall), 0.0, None)
        if not self._zero_boundary:  # the outer halves of the first / last function lie outside the grid
            vals = np.where((x < self._domain[0]) | (x > self._domain[1]), 0.0, vals)
        return vals

    def _evaluate(self, x):
        return self._hats(x)

    def eval_elem(self, idx: int, x):
        idx = int(idx) % len(self)
        return self._hats(x, slice(idx, idx + 1))[..., 0]

    def support_bounds(self, idx: int):
        m = len(self)
        if not -m <= idx < m:
            raise IndexError(idx)
        idx %= m
        lo, hi = self._grid[idx], self._grid[idx + 2]
        if not self._zero_boundary:
            lo, hi = max(lo, self._domain[0]), min(hi, self._domain[1])
        return lo, hi

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
