"""L2 projection onto a piecewise-linear basis: ``f -> M^{-1} (int phi_i f)_i`` with the mass matrix ``M`` of the basis.

``L2Projection_UnivariateLinearInterpolationBasis`` follows src/linpde_gp/linfunctls/projections/l2/_fem.py:14-95
(``normalizer`` :38-62, functions :64-95) and, applied to covariance functions and process-vector cross-covariances,
src/linpde_gp/randprocs/covfuncs/linfunctls/_registry.py:198-238 and
src/linpde_gp/randprocs/crosscov/linfunctls/projections.py:18-122.  The reference evaluates ``P k(x, .)`` in closed form
for Matern-3/2 only and integrates every other case with ``scipy.integrate.quad`` / ``dblquad`` point by point; here

* half-integer Matern kernels of any order use the closed-form device kernel ``lpgp_matern_hat_integral``;
* kernels that are smooth inside the elements (ExpQuad, and the second stage ``P k P*`` of every kernel) use
  Gauss-Legendre quadrature element by element, evaluated as ONE pairwise Gram launch at the quadrature nodes followed
  by a DMMA GEMM with the (basis x nodes) weight matrix;
* the normaliser ``M^{-1}`` is applied with the device Cholesky factor of the tridiagonal mass matrix.

As an observation functional (``condition_on_observations(Y, L=proj)``) it contributes one atom of kind ``"proj"``."""
from __future__ import annotations

import numpy as np

from .. import LinearFunctional
from ...functions import Constant, Function
from ...functions.bases import UnivariateLinearInterpolationBasis


class L2Projection_UnivariateLinearInterpolationBasis(LinearFunctional):  # pylint: disable=invalid-name
    def __init__(self, basis: UnivariateLinearInterpolationBasis, *, normalized: bool = True) -> None:
        if not isinstance(basis, UnivariateLinearInterpolationBasis):
            raise TypeError("`basis` must be a `UnivariateLinearInterpolationBasis`")
        self._basis = basis
        self._normalized = bool(normalized)
        self._mass_factor = None
        self._quad = {}
        super().__init__(input_shapes=((), ()), output_shape=basis.output_shape)

    basis = property(lambda self: self._basis)
    normalized = property(lambda self: self._normalized)

    # -- mass matrix / normaliser (_fem.py:38-62) -----------------------------------------------------------------
    def mass_matrix(self) -> np.ndarray:
        """``M_ij = int phi_i phi_j``, assembled element by element: an element of length h contributes h/3 to the two
        diagonal entries of its end nodes and h/6 to their coupling (the tridiagonal matrix of _fem.py:38-62)."""
        b = self._basis
        e = b.elements()
        h = np.diff(e)
        first = 1 if b.zero_boundary else 0  # index (among the element nodes) of the first node that carries a function
        m = len(b)
        M = np.zeros((m, m))
        for el, hl in enumerate(h):
            i, j = el - first, el + 1 - first  # basis functions of the element's left / right node
            if 0 <= i < m:
                M[i, i] += hl / 3.0
            if 0 <= j < m:
                M[j, j] += hl / 3.0
            if 0 <= i < m and 0 <= j < m:
                M[i, j] += hl / 6.0
                M[j, i] += hl / 6.0
        return M

    def _device_mass_factor(self):
        """Device Cholesky factor of the mass matrix (SPD, tridiagonal; factored densely -- m x m is small)."""
        from ... import backend  # pylint: disable=import-outside-toplevel

        if self._mass_factor is None:
            m = len(self._basis)
            f = backend.DeviceFactor([m + (m % 2)])
            f.L.zero_()
            f.L[:m, :m].copy_(backend.to_device(self.mass_matrix()))
            if m % 2:
                f.L[m, m] = 1.0
            f.potrf()
            self._mass_factor = f
        return self._mass_factor

    def normalize_rows(self, R):
        """``R <- R M^{-1}`` in place for a device matrix whose ROWS are un-normalised projections (no-op if the
        projection is not normalised)."""
        if not self._normalized:
            return R
        from ... import backend  # pylint: disable=import-outside-toplevel

        f = self._device_mass_factor()
        m = len(self._basis)
        if f.n == m and R.stride(0) % 2 == 0 and R.data_ptr() % 16 == 0:
            f.potrs(R)
            return R
        T = backend.alloc_matrix(R.shape[0], f.n).zero_()
        T[:, :m].copy_(R)
        f.potrs(T)
        R.copy_(T[:, :m])
        return R

    def normalizer(self, v, axis: int = -1) -> np.ndarray:
        """Host-array version (prior-mean projections): ``M^{-1}`` applied along ``axis``."""
        if not self._normalized:
            return np.asarray(v, dtype=np.double)
        from ... import backend  # pylint: disable=import-outside-toplevel

        v = np.moveaxis(np.asarray(v, dtype=np.double), axis, -1)
        R = backend.alloc_matrix(int(np.prod(v.shape[:-1], dtype=np.int64)) if v.ndim > 1 else 1, v.shape[-1])
        R.copy_(backend.to_device(np.ascontiguousarray(v.reshape(R.shape))))
        self.normalize_rows(R)
        return np.moveaxis(R.cpu().numpy().reshape(v.shape), -1, axis)

    # -- quadrature data on the device --------------------------------------------------------------------------
    def device_quadrature(self, order: int = 20):
        """``(T, W)``: Gauss-Legendre nodes as device points (Q x 1) and the device weight matrix (m x Q)."""
        from ... import backend  # pylint: disable=import-outside-toplevel

        if order not in self._quad:
            nodes, W = self._basis.gauss_legendre(order)
            Wd = backend.alloc_matrix(*W.shape)
            Wd.copy_(backend.to_device(W))
            self._quad[order] = (backend.points(nodes, 1), Wd, nodes)
        return self._quad[order]

    def device_grid(self):
        from ... import backend  # pylint: disable=import-outside-toplevel

        if "grid" not in self._quad:
            self._quad["grid"] = backend.to_device(self._basis.grid)
        return self._quad["grid"]

    # -- application ---------------------------------------------------------------------------------------------------
    def __call__(self, f, /, **kwargs):
        if isinstance(f, Constant) and f.output_shape == ():  # _fem.py:83-95
            b = self._basis
            res = (float(f.value) / 2.0) * (b.x_ip1 - b.x_im1)
            if not b.zero_boundary:
                res[0] = (float(f.value) / 2.0) * (b.x_ip1[0] - b.x_i[0])
                res[-1] = (float(f.value) / 2.0) * (b.x_i[-1] - b.x_im1[-1])
            return self.normalizer(res, axis=-1)
        if isinstance(f, Function):  # _fem.py:64-81 (scipy.integrate.quad there; Gauss-Legendre per element here)
            if f.input_shape != () or f.output_shape != ():
                raise NotImplementedError("L2 projections of scalar functions on the real line only")
            nodes, W = self._basis.gauss_legendre(32)
            return self.normalizer(W @ np.asarray(f(nodes), dtype=np.double), axis=-1)
        from ...randprocs import covfuncs, crosscov  # pylint: disable=import-outside-toplevel

        if isinstance(f, covfuncs.CovarianceFunction) and not isinstance(f, covfuncs.Zero):
            # dispatch of covfuncs/linfunctls/_registry.py:198-238
            argnum = kwargs.get("argnum", 0)
            if argnum not in (0, 1):
                raise ValueError("`argnum` must either be 0 or 1.")
            if f.input_shape != ():
                raise ValueError("L2 projections onto a univariate basis need a process on the real line")
            cls = (crosscov.Matern32_L2Projection_UnivariateLinearInterpolationBasis
                   if isinstance(f, covfuncs.Matern) and f.nu == 1.5
                   else crosscov.CovarianceFunction_L2Projection_UnivariateLinearInterpolationBasis)
            return cls(f, self, reverse=(argnum == 0))
        return super().__call__(f, **kwargs)

    def _atoms(self):
        return [(1.0, "proj", None, self)]
