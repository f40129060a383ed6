"""Device-resident linear operators of the hot path.

Mirrors the slice of ``probnum.linops`` / ``linpde_gp.linops`` the conditioning path touches:
``CovarianceLinearOperator`` (pn/randprocs/covfuncs/_covariance_linear_operator.py:22-90: lazy handle that
densifies on demand), ``LinearOperator.todense/cholesky/solve/inv`` with cached factorisations
(pn/linops/_linear_operator.py:221-315, 384-410, 784-865) and the bordered block factor of
``BlockMatrix2x2`` (src/linpde_gp/linops/_block.py:84-292).  Matrices live in HBM as torch tensors; every
operation runs through ``liblpgp.so``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import backend


class LinearOperator:
    def __init__(self, shape):
        self._shape = (int(shape[0]), int(shape[1]))
        self.is_symmetric: Optional[bool] = None
        self.is_positive_definite: Optional[bool] = None
        self.is_lower_triangular: Optional[bool] = None
        self.is_upper_triangular: Optional[bool] = None

    @property
    def shape(self):
        return self._shape

    @property
    def dtype(self):
        return np.dtype(np.double)

    @property
    def ndim(self):
        return 2

    @property
    def is_square(self):
        return self._shape[0] == self._shape[1]

    def device_dense(self) -> torch.Tensor:  # pragma: no cover - abstract
        raise NotImplementedError

    def todense(self, cache: bool = True) -> np.ndarray:
        return self.device_dense()[: self.shape[0], : self.shape[1]].cpu().numpy()

    @property
    def T(self):
        return _Transposed(self)

    def __matmul__(self, other):
        if isinstance(other, LinearOperator):
            other = other.todense()
        x = np.asarray(other, dtype=np.double)
        vec = x.ndim == 1
        if vec:
            x = x[:, None]
        if x.shape[0] != self.shape[1]:
            raise ValueError(f"shape mismatch: {self.shape} @ {x.shape}")
        A = self.device_dense()
        Bt = backend.alloc_matrix(x.shape[1], x.shape[0])
        Bt.copy_(backend.to_device(np.ascontiguousarray(x.T)))
        if A.stride(0) % 2 or A.data_ptr() % 16:
            A2 = backend.alloc_matrix(*A.shape)
            A2.copy_(A)
            A = A2
        C = backend.alloc_matrix(self.shape[0], x.shape[1])
        backend.gemm_nt(A, Bt, C, 1.0, 0.0)
        res = C.cpu().numpy()
        return res[:, 0] if vec else res


class _Transposed(LinearOperator):
    def __init__(self, op):
        super().__init__((op.shape[1], op.shape[0]))
        self._op = op
        self.is_symmetric = op.is_symmetric

    @property
    def T(self):
        return self._op

    def device_dense(self):
        return self._op.device_dense().T.contiguous()


class Matrix(LinearOperator):
    """Dense matrix given on the host, kept on the device (``pn.linops.Matrix``)."""

    def __init__(self, A):
        A = np.asarray(A, dtype=np.double)
        if A.ndim != 2:
            raise ValueError("2-D array expected")
        super().__init__(A.shape)
        self._dev = backend.alloc_matrix(*A.shape)
        self._dev.copy_(backend.to_device(A))

    def device_dense(self):
        return self._dev


class Scaling(LinearOperator):
    """Diagonal operator ``diag(factors)`` (``pn.linops.Scaling``): the usual observation-noise covariance."""

    def __init__(self, factors, shape=None):
        factors = np.asarray(factors, dtype=np.double)
        if factors.ndim == 0:
            if shape is None:
                raise ValueError("scalar Scaling needs a shape")
            n = int(shape[0]) if np.ndim(shape) else int(shape)
            factors = np.full(n, float(factors))
        super().__init__((factors.size, factors.size))
        self.factors = factors.reshape(-1)
        self.is_symmetric = True

    def device_dense(self):
        return torch.diag(backend.to_device(self.factors))

    def todense(self, cache=True):
        return np.diag(self.factors)


class CovarianceLinearOperator(LinearOperator):
    """Lazy covariance matrix ``k(x0, x1)``; nothing is computed until it is densified / factorised."""

    def __init__(self, covfunc, x0: np.ndarray, x1: Optional[np.ndarray]):
        n0 = x0.shape[0]
        n1 = n0 if x1 is None else x1.shape[0]
        super().__init__((n0, n1))
        self._covfunc = covfunc
        self._x0 = x0
        self._x1 = x1
        self._dense: Optional[torch.Tensor] = None
        self._factor = None
        if x1 is None:
            self.is_symmetric = True

    @property
    def covfunc(self):
        return self._covfunc

    def assemble_into(self, out: torch.Tensor, lower: bool = False, X0=None, X1=None, accumulate: bool = False) -> torch.Tensor:
        """Write (or add, ``accumulate=True``) the block into ``out`` (a view into a larger device buffer) with the
        Gram kernel."""
        from .randprocs import covfuncs

        d = self._covfunc.input_size
        X0 = backend.points(self._x0, d) if X0 is None else X0
        if self._x1 is not None and X1 is None:
            X1 = backend.points(self._x1, d)
        k = self._covfunc
        if isinstance(k, covfuncs.SumCovarianceFunction):
            try:
                descs = [k.descriptor()]
            except NotImplementedError:
                descs = k.descriptors()
        else:
            descs = [k.descriptor()]
        for i, desc in enumerate(descs):
            backend.gram(desc, X0, X1, out=out, lower=lower, accumulate=accumulate or i > 0)
        return out

    def device_dense(self) -> torch.Tensor:
        if self._dense is None:
            out = backend.alloc_matrix(*self.shape)
            self.assemble_into(out)
            self._dense = out
        return self._dense

    # -- SPD solves (pn/linops/_linear_operator.py:267-315) ------------------------------------------------
    def cholesky(self, lower: bool = True) -> "CholeskyFactor":
        if not self.is_square:
            raise np.linalg.LinAlgError("The Cholesky decomposition is only defined for square matrices.")
        if self.is_symmetric is False:
            raise np.linalg.LinAlgError("The Cholesky decomposition is only defined for symmetric matrices.")
        if self._factor is None:
            n = self.shape[0]
            f = backend.DeviceFactor([n + n % 2])
            self.assemble_into(f.L[:n, :n], lower=True)
            if n % 2:
                f.L[n, : n + 1] = 0.0
                f.L[n, n] = 1.0
            try:
                f.potrf()
            except np.linalg.LinAlgError:
                self.is_positive_definite = False
                raise
            self.is_positive_definite = True
            self._factor = CholeskyFactor(f, n)
        return self._factor if lower else self._factor.T

    def solve(self, B):
        return self.cholesky(True).solve_spd(B)


class CholeskyFactor(LinearOperator):
    """Lower-triangular factor ``L`` of an SPD matrix, resident on the device together with its inverted
    diagonal blocks.  ``solve_spd`` solves with ``L L^T``."""

    def __init__(self, factor: backend.DeviceFactor, n: int):
        super().__init__((n, n))
        self.factor = factor
        self.is_lower_triangular = True

    def device_dense(self):
        n = self.shape[0]
        return torch.tril(self.factor.L[:n, :n])

    def solve_spd(self, B):
        """``(L L^T)^{-1} B`` for a vector, a matrix of column right-hand sides or a stack (..., n, k)."""
        B = np.asarray(B, dtype=np.double)
        n, nphys = self.shape[0], self.factor.n
        if B.ndim == 1:
            if B.shape[0] != n:
                raise ValueError("`b` has the wrong length")
            rows = B[None, :]
        elif B.ndim >= 2:
            if B.shape[-2] != n:
                raise ValueError("`b` must be a vector or a (stack of) matrices.")
            rows = np.moveaxis(B, -2, -1).reshape(-1, n)
        dev = backend.alloc_matrix(rows.shape[0], nphys)
        dev.zero_()
        dev[:, :n].copy_(backend.to_device(np.ascontiguousarray(rows)))
        self.factor.potrs(dev)  # forward + backward substitution per right-hand side
        res = dev[:, :n].cpu().numpy()
        if B.ndim == 1:
            return res[0]
        return np.moveaxis(res.reshape(B.shape[:-2] + (B.shape[-1], n)), -1, -2)

    @property
    def T(self):
        t = _Transposed(self)
        t.is_upper_triangular = True
        return t
