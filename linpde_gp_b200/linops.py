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

    def assemble_into(self, out: torch.Tensor, lower: bool = False, accumulate: bool = False) -> torch.Tensor:
        """Write (or add) the matrix into ``out`` (a view into a larger device buffer); structured operators
        override this with their own assembly kernel."""
        D = self.device_dense()[: self.shape[0], : self.shape[1]]
        if accumulate:
            out.add_(D)
        else:
            out.copy_(D)
        return out

    def kron_terms(self):
        """``[(alpha, A_dev, B_dev), ...]`` with ``self == sum alpha * kron(A, B)``, or ``None``."""
        return None

    # -- arithmetic (pn/linops/_arithmetic_fallbacks.py: ScaledLinearOperator / SumLinearOperator) -------------
    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledLinearOperator(self, other)
        return NotImplemented

    def __neg__(self):
        return ScaledLinearOperator(self, -1.0)

    def __add__(self, other):
        if isinstance(other, LinearOperator):
            return SumLinearOperator(self, other)
        return NotImplemented

    def __radd__(self, other):
        if np.ndim(other) == 0 and other == 0:  # sum([...]) / the reference's `res_zero_value=0`
            return self
        return NotImplemented

    # -- SPD solves (pn/linops/_linear_operator.py:267-315) ------------------------------------------------
    _factor = None

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


def _aligned(A: torch.Tensor) -> torch.Tensor:
    """``A`` itself if the DMMA GEMM can read it through TMA (even leading dimension, 16-byte aligned), else a copy."""
    if A.dim() == 2 and A.stride(1) == 1 and A.stride(0) % 2 == 0 and A.data_ptr() % 16 == 0 and A.stride(0) >= A.shape[1]:
        return A
    B = backend.alloc_matrix(*A.shape)
    B.copy_(A)
    return B


class ScaledLinearOperator(LinearOperator):
    """``scalar * A`` (pn/linops/_arithmetic_fallbacks.py:26-60)."""

    def __init__(self, linop: LinearOperator, scalar):
        if np.ndim(scalar) != 0:
            raise TypeError("`scalar` must be a scalar")
        super().__init__(linop.shape)
        self._linop = linop
        self._scalar = float(scalar)
        self.is_symmetric = linop.is_symmetric

    def kron_terms(self):
        t = self._linop.kron_terms()
        return None if t is None else [(self._scalar * a, A, B) for a, A, B in t]

    def device_dense(self):
        t = self.kron_terms()
        if t is not None:
            return backend.kron_sum(t)
        return self._scalar * self._linop.device_dense()[: self.shape[0], : self.shape[1]]

    def assemble_into(self, out, lower=False, accumulate=False):
        t = self.kron_terms()
        if t is not None:
            return backend.kron_sum(t, out=out, lower=lower, accumulate=accumulate)
        return super().assemble_into(out, lower=lower, accumulate=accumulate)

    def __matmul__(self, other):
        return self._scalar * (self._linop @ other)


class SumLinearOperator(LinearOperator):
    """``A + B + ...`` (pn/linops/_arithmetic_fallbacks.py:63-117); nested sums are flattened."""

    def __init__(self, *summands: LinearOperator):
        flat = []
        for s_ in summands:
            flat.extend(s_._summands if isinstance(s_, SumLinearOperator) else [s_])
        if not all(s_.shape == flat[0].shape for s_ in flat):
            raise ValueError("All summands must have the same shape.")
        super().__init__(flat[0].shape)
        self._summands = tuple(flat)
        if all(s_.is_symmetric for s_ in flat):
            self.is_symmetric = True

    def kron_terms(self):
        out = []
        for s_ in self._summands:
            t = s_.kron_terms()
            if t is None:
                return None
            out.extend(t)
        if not all(A.shape == out[0][1].shape and B.shape == out[0][2].shape for _, A, B in out):
            return None
        return out

    def assemble_into(self, out, lower=False, accumulate=False):
        t = self.kron_terms()
        if t is not None:
            return backend.kron_sum(t, out=out, lower=lower, accumulate=accumulate)
        for i, s_ in enumerate(self._summands):
            s_.assemble_into(out, lower=lower, accumulate=accumulate or i > 0)
        return out

    def device_dense(self):
        out = backend.alloc_matrix(*self.shape)
        return self.assemble_into(out)

    def __matmul__(self, other):
        res = None
        for s_ in self._summands:
            v = s_ @ other
            res = v if res is None else res + v
        return res


class Kronecker(LinearOperator):
    """Kronecker product ``A (x) B`` (pn/linops/_kronecker.py:17-166) with device-resident factors.

    ``todense`` / ``assemble_into`` run the Kronecker assembly kernel (``lpgp_kron_sum``, one multiply per entry,
    HBM-write bound); ``@`` uses ``(A (x) B) vec(X) = vec(A X B^T)`` -- two DMMA GEMMs with the small factors, the
    big matrix is never formed."""

    def __init__(self, A: LinearOperator, B: LinearOperator):
        self.A = A if isinstance(A, LinearOperator) else Matrix(A)
        self.B = B if isinstance(B, LinearOperator) else Matrix(B)
        super().__init__((self.A.shape[0] * self.B.shape[0], self.A.shape[1] * self.B.shape[1]))
        if self.A.is_symmetric and self.B.is_symmetric:
            self.is_symmetric = True
        self._mats = None

    def _factor_matrices(self):
        if self._mats is None:  # the (small) factor matrices are assembled once
            self._mats = (self.A.device_dense()[: self.A.shape[0], : self.A.shape[1]],
                          self.B.device_dense()[: self.B.shape[0], : self.B.shape[1]])
        return self._mats

    def kron_terms(self):
        A, B = self._factor_matrices()
        return [(1.0, A, B)]

    def device_dense(self):
        return backend.kron_sum(self.kron_terms())

    def assemble_into(self, out, lower=False, accumulate=False):
        return backend.kron_sum(self.kron_terms(), out=out, lower=lower, accumulate=accumulate)

    def __matmul__(self, other):
        if isinstance(other, LinearOperator):
            other = other.todense()
        x = np.asarray(other, dtype=np.double)
        vec = x.ndim == 1
        if vec:
            x = x[:, None]
        if x.shape[0] != self.shape[1]:
            raise ValueError(f"shape mismatch: {self.shape} @ {x.shape}")
        A, B = self._factor_matrices()
        (n1, m1), (n2, m2) = A.shape, B.shape
        k = x.shape[1]
        Xd = backend.to_device(np.ascontiguousarray(x.T)).reshape(k * m1, m2)  # rows (c, j1), columns j2
        T = backend.alloc_matrix(k * m1, n2)
        backend.gemm_nt(_aligned(Xd), _aligned(B), T, 1.0, 0.0)                # T[(c, j1), i2] = sum_j2 X B[i2, j2]
        T2 = _aligned(T.reshape(k, m1, n2).permute(0, 2, 1).reshape(k * n2, m1))  # rows (c, i2), columns j1
        U = backend.alloc_matrix(k * n2, n1)
        backend.gemm_nt(T2, _aligned(A), U, 1.0, 0.0)                           # U[(c, i2), i1] = sum_j1 T A[i1, j1]
        res = U.reshape(k, n2, n1).permute(2, 1, 0).reshape(n1 * n2, k).cpu().numpy()
        return res[:, 0] if vec else res


class BlockMatrix(LinearOperator):
    """Dense block matrix ``[[A_00, A_01, ...], [A_10, ...], ...]`` of linear operators
    (src/linpde_gp/linops/_block.py:17-80); blocks are assembled straight into their place in one device buffer."""

    def __init__(self, blocks):
        self._blocks = [list(row) for row in blocks]
        if not self._blocks or any(len(row) != len(self._blocks[0]) for row in self._blocks):
            raise ValueError("blocks must form a rectangular grid")
        self._row_sizes = [row[0].shape[0] for row in self._blocks]
        self._col_sizes = [b.shape[1] for b in self._blocks[0]]
        for row, rs in zip(self._blocks, self._row_sizes):
            for b, cs in zip(row, self._col_sizes):
                if b.shape != (rs, cs):
                    raise ValueError("inconsistent block shapes")
        super().__init__((sum(self._row_sizes), sum(self._col_sizes)))

    @property
    def blocks(self):
        return self._blocks

    def device_dense(self):
        out = backend.alloc_matrix(*self.shape)
        r = 0
        for row, rs in zip(self._blocks, self._row_sizes):
            c = 0
            for b, cs in zip(row, self._col_sizes):
                b.assemble_into(out[r : r + rs, c : c + cs])
                c += cs
            r += rs
        return out


class BlockDiagonalMatrix(LinearOperator):
    """``diag(A_0, A_1, ...)`` with possibly non-square blocks (pn/linops/_block.py ``BlockDiagonalMatrix``): the
    covariance matrix of a process with independent outputs."""

    def __init__(self, *blocks: LinearOperator):
        self._blocks = tuple(blocks)
        super().__init__((sum(b.shape[0] for b in blocks), sum(b.shape[1] for b in blocks)))
        if all(b.is_symmetric for b in blocks):
            self.is_symmetric = True

    @property
    def blocks(self):
        return self._blocks

    def device_dense(self):
        out = backend.alloc_matrix(*self.shape)
        out.zero_()
        r = c = 0
        for b in self._blocks:
            b.assemble_into(out[r : r + b.shape[0], c : c + b.shape[1]])
            r += b.shape[0]
            c += b.shape[1]
        return out
