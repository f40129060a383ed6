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
        if not self.is_symmetric:  # None (unknown) raises as well, like pn/linops/_linear_operator.py:823-826
            raise np.linalg.LinAlgError("The Cholesky decomposition is only defined for symmetric matrices.")
        if self.is_positive_definite is False:
            raise np.linalg.LinAlgError("The linear operator is not positive definite.")
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

    def _adopt_factor(self, fac: "CholeskyFactor") -> None:
        self._factor = fac
        self.is_positive_definite = True

    def solve(self, B):
        """``A^{-1} B`` (pn/linops/_linear_operator.py:221-315): triangular operators substitute, symmetric ones go
        through the cached Cholesky factor.  There is no LU kernel on this path (the reference's last-resort branch
        :311-315): an operator that is not flagged symmetric (``is_symmetric`` None or False) is refused instead of
        being factored from its lower triangle."""
        if self.is_lower_triangular or self.is_upper_triangular:
            return self._triangular_solve(B)
        if not self.is_symmetric:
            raise NotImplementedError(
                "solve() of an operator that is not flagged symmetric needs an LU factorisation, which the device path "
                "does not provide; set `is_symmetric = True` if the matrix is symmetric positive definite")
        return self.cholesky(True).solve_spd(B)

    def _triangular_solve(self, B):
        raise NotImplementedError(f"{type(self).__name__} has no triangular solve")

    def _solve_rows_device(self, R: torch.Tensor) -> torch.Tensor:
        """Device matrix whose rows are ``A^{-1} r`` for the rows ``r`` of the device matrix ``R`` (building block of the
        structured Kronecker solve): symmetric operators through their cached Cholesky factor."""
        if not self.is_symmetric:
            raise NotImplementedError(f"row solves with {type(self).__name__} (not flagged symmetric)")
        return self.cholesky(True)._spd_rows_device(R)  # pylint: disable=protected-access

    def inv(self) -> "LinearOperator":
        """Lazy inverse (pn/linops/_linear_operator.py:1009-1028, 1474-1521): ``inv() @ B`` is ``solve(B)``."""
        if not self.is_square:
            raise np.linalg.LinAlgError("Only square operators can be inverted.")
        return _InverseLinearOperator(self)

    def logabsdet(self) -> float:
        if self.is_lower_triangular or self.is_upper_triangular:
            d = torch.diagonal(self.device_dense()[: self.shape[0], : self.shape[1]])
            return float(torch.log(torch.abs(d)).sum().item())
        return self.cholesky(True).factor.logdet()

    def det(self) -> float:
        if self.is_lower_triangular or self.is_upper_triangular:
            d = torch.diagonal(self.device_dense()[: self.shape[0], : self.shape[1]])
            return float(torch.prod(d).item())
        return float(np.exp(self.logabsdet()))

    def trace(self) -> float:
        if not self.is_square:
            raise ValueError("The trace is only defined for square operators.")
        return float(torch.diagonal(self.device_dense()[: self.shape[0], : self.shape[1]]).sum().item())

    @property
    def T(self):
        if self.is_symmetric:
            return self
        return _Transposed(self)

    def __matmul__(self, other):
        if hasattr(other, "input_shapes") and hasattr(other, "_atoms"):  # LinearFunctional: its __rmatmul__ builds A @ l
            return NotImplemented
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
        self.is_positive_definite = op.is_positive_definite
        self.is_lower_triangular = op.is_upper_triangular
        self.is_upper_triangular = op.is_lower_triangular

    @property
    def T(self):
        return self._op

    def device_dense(self):
        return self._op.device_dense()[: self._op.shape[0], : self._op.shape[1]].T.contiguous()

    def _triangular_solve(self, B):
        return self._op._triangular_solve(B, trans=True)  # pylint: disable=protected-access

    def _solve_rows_device(self, R):
        if isinstance(self._op, CholeskyFactor):
            return self._op._solve_rows_device(R, trans=True)  # pylint: disable=protected-access
        return super()._solve_rows_device(R)


class _InverseLinearOperator(LinearOperator):
    """``A^{-1}`` applied by solving with ``A`` (pn/linops/_linear_operator.py:1474-1521)."""

    def __init__(self, op: LinearOperator):
        super().__init__(op.shape)
        self._op = op
        self.is_symmetric = op.is_symmetric
        self.is_positive_definite = op.is_positive_definite
        self.is_lower_triangular = op.is_lower_triangular
        self.is_upper_triangular = op.is_upper_triangular

    def inv(self):
        return self._op

    @property
    def T(self):
        return self if self.is_symmetric else _InverseLinearOperator(self._op.T)

    def __matmul__(self, other):
        if isinstance(other, LinearOperator):
            other = other.todense()
        return self._op.solve(other)

    def solve(self, B):
        return self._op @ B

    def todense(self, cache: bool = True) -> np.ndarray:
        return self._op.solve(np.eye(self.shape[0]))

    def device_dense(self):
        return backend.to_device(self.todense())

    def det(self):
        return 1.0 / self._op.det()

    def logabsdet(self):
        return -self._op.logabsdet()


def aslinop(A) -> LinearOperator:
    """``pn.linops.aslinop``: linear operators pass through, arrays become device-resident ``Matrix`` objects."""
    return A if isinstance(A, LinearOperator) else Matrix(np.atleast_2d(np.asarray(A, dtype=np.double)))


def _rows_of(B, n):
    """Right-hand sides ``(n,)``, ``(n, k)`` or ``(..., n, k)`` as an array of rows ``(nrhs, n)`` + the inverse map."""
    B = np.asarray(B, dtype=np.double)
    if B.ndim == 1:
        if B.shape[0] != n:
            raise ValueError("`b` has the wrong length")
        return B[None, :], (lambda res: res[0])
    if B.shape[-2] != n:
        raise ValueError("`b` must be a vector or a (stack of) matrices.")
    rows = np.moveaxis(B, -2, -1).reshape(-1, n)
    return rows, (lambda res: np.moveaxis(res.reshape(B.shape[:-2] + (B.shape[-1], n)), -1, -2))


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


class Identity(Scaling):
    """``pn.linops.Identity``."""

    def __init__(self, shape):
        n = int(shape[0]) if np.ndim(shape) else int(shape)
        if np.ndim(shape) and int(shape[0]) != int(shape[-1]):
            raise ValueError("Identity must be square")
        super().__init__(np.ones(n))
        self.is_positive_definite = True
        self.is_lower_triangular = self.is_upper_triangular = True

    def _triangular_solve(self, B, trans: bool = False):
        return np.array(B, dtype=np.double)


class Zero(LinearOperator):
    """``pn.linops.Zero``: the all-zero block of a block-diagonal / block-triangular matrix."""

    def __init__(self, shape):
        super().__init__(shape)
        if self.is_square:
            self.is_symmetric = True

    @property
    def T(self):
        return Zero((self.shape[1], self.shape[0]))

    def device_dense(self):
        out = backend.alloc_matrix(*self.shape)
        return out.zero_()

    def assemble_into(self, out, lower=False, accumulate=False):
        return out if accumulate else out.zero_()

    def todense(self, cache=True):
        return np.zeros(self.shape)

    def __matmul__(self, other):
        if isinstance(other, LinearOperator):
            other = other.todense()
        x = np.asarray(other, dtype=np.double)
        if x.shape[0 if x.ndim == 1 else -2] != self.shape[1]:
            raise ValueError(f"shape mismatch: {self.shape} @ {x.shape}")
        return np.zeros((self.shape[0],) if x.ndim == 1 else x.shape[:-2] + (self.shape[0], x.shape[-1]))


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
        descs = covfuncs.device_descriptors(self._covfunc)  # scalars distributed over sums, zero summands dropped
        if not descs and not accumulate:
            out.zero_()
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
    diagonal blocks.  ``solve_spd`` solves with ``L L^T``; ``solve`` / ``inv() @`` substitute with ``L`` itself and
    ``.T.solve`` with ``L^T`` (scipy ``solve_triangular``, pn/linops/_linear_operator.py:296-299).

    ``index`` maps the logical rows to the rows of the device factor: bordered factors pad every observation batch
    to an even size with an identity row, so the logical matrix is a sub-matrix of the physical one."""

    def __init__(self, factor: backend.DeviceFactor, n=None, index=None):
        if index is None:
            index = np.arange(int(n))
        index = np.asarray(index, dtype=np.int64)
        super().__init__((len(index), len(index)))
        self.factor = factor
        self.index = index
        self.is_lower_triangular = True

    def _dev_index(self):
        return torch.as_tensor(self.index, device=self.factor.L.device)

    def device_dense(self):
        n = self.shape[0]
        if n == self.factor.n or np.array_equal(self.index, np.arange(n)):
            return torch.tril(self.factor.L[:n, :n])
        idx = self._dev_index()
        return torch.tril(self.factor.L)[idx][:, idx].contiguous()

    def _scatter(self, rows: np.ndarray) -> torch.Tensor:
        dev = backend.alloc_matrix(rows.shape[0], self.factor.n)
        dev.zero_()
        dev[:, self._dev_index()] = backend.to_device(np.ascontiguousarray(rows))
        return dev

    def solve_spd(self, B):
        """``(L L^T)^{-1} B`` for a vector, a matrix of column right-hand sides or a stack (..., n, k)."""
        rows, back = _rows_of(B, self.shape[0])
        dev = self._scatter(rows)
        self.factor.potrs(dev)  # forward + backward substitution per right-hand side
        return back(dev[:, self._dev_index()].cpu().numpy())

    def _apply_rows_device(self, R: torch.Tensor, how: str) -> torch.Tensor:
        """Rows of the device matrix ``R`` (logical width) through ``potrs`` / ``L^{-1}`` / ``L^{-T}`` of the physical
        factor (zero-padded columns for the identity padding rows); returns a device matrix of the logical width."""
        n = self.shape[0]
        direct = n == self.factor.n and R.stride(0) % 2 == 0 and R.data_ptr() % 16 == 0 and R.stride(1) == 1
        if direct:
            dev = R
        else:
            dev = backend.alloc_matrix(R.shape[0], self.factor.n).zero_()
            dev[:, self._dev_index()] = R
        if how == "spd":
            self.factor.potrs(dev)
        elif how == "fwd":
            self.factor.trsm_rlt(dev)  # rows <- rows L^{-T}, i.e. L^{-1} b for every right-hand side b
        elif dev.shape[0] >= 4:
            self.factor.trsm_rln(dev)  # rows <- rows L^{-1}, i.e. L^{-T} b: blocked DMMA solve
        else:
            for r in range(dev.shape[0]):
                self.factor.trsv(dev[r], trans=True)
        return dev if direct else dev[:, self._dev_index()]

    def _spd_rows_device(self, R: torch.Tensor) -> torch.Tensor:
        return self._apply_rows_device(R, "spd")

    def _solve_rows_device(self, R: torch.Tensor, trans: bool = False) -> torch.Tensor:
        return self._apply_rows_device(R, "bwd" if trans else "fwd")

    def _triangular_solve(self, B, trans: bool = False):
        """``L^{-1} B`` (``trans``: ``L^{-T} B``).  The padding rows of the physical factor are identity rows, so
        substituting with the physical factor on zero-padded right-hand sides gives the logical solution."""
        rows, back = _rows_of(B, self.shape[0])
        R = backend.alloc_matrix(*rows.shape)
        R.copy_(backend.to_device(np.ascontiguousarray(rows)))
        return back(self._apply_rows_device(R, "bwd" if trans else "fwd").cpu().numpy())

    def logabsdet(self) -> float:
        return 0.5 * self.factor.logdet()

    def det(self) -> float:
        return float(np.exp(self.logabsdet()))

    @property
    def T(self):
        return _Transposed(self)


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

    # a scalar multiple of a structured (Kronecker) operator keeps its structure: (c A)^{-1} = A^{-1} / c
    def solve(self, B):
        if isinstance(self._linop, Kronecker):
            return self._linop.solve(B) / self._scalar
        return super().solve(B)

    def logabsdet(self) -> float:
        if isinstance(self._linop, Kronecker):
            return self._linop.logabsdet() + self.shape[0] * float(np.log(abs(self._scalar)))
        return super().logabsdet()


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

    # -- structure-preserving algebra (pn/linops/_kronecker.py:122-166, 233-242): nothing of size N x N is formed -----
    @property
    def T(self):
        return self if self.is_symmetric else Kronecker(self.A.T, self.B.T)

    def cholesky(self, lower: bool = True) -> "Kronecker":
        """``chol(A (x) B) = chol(A) (x) chol(B)`` (_kronecker.py:233-242)."""
        if not (self.A.is_symmetric and self.B.is_symmetric):
            raise np.linalg.LinAlgError("The Cholesky decomposition is only defined for symmetric matrices.")
        K = Kronecker(self.A.cholesky(lower), self.B.cholesky(lower))
        K.is_lower_triangular, K.is_upper_triangular = bool(lower), not lower
        return K

    def inv(self) -> "Kronecker":
        """``(A (x) B)^{-1} = A^{-1} (x) B^{-1}`` (_kronecker.py:135-140)."""
        if not (self.A.is_square and self.B.is_square):
            raise np.linalg.LinAlgError("Only square operators can be inverted.")
        return Kronecker(self.A.inv(), self.B.inv())

    def solve(self, B):
        """``(A (x) B)^{-1} vec(X) = vec(A^{-1} X B^{-T})``: two multi-right-hand-side solves with the SMALL factors (their
        cached Cholesky factors, blocked DMMA substitutions), O(n1 n2 (n1 + n2)) instead of O((n1 n2)^3)."""
        if not (self.A.is_square and self.B.is_square):
            raise np.linalg.LinAlgError("Only square operators can be solved with.")
        rows, back = _rows_of(B, self.shape[0])
        return back(self._solve_rows_device(backend.to_device(np.ascontiguousarray(rows))).cpu().numpy())

    def _solve_rows_device(self, R: torch.Tensor) -> torch.Tensor:
        n1, n2 = self.A.shape[0], self.B.shape[0]
        k = R.shape[0]
        X = backend.alloc_matrix(k * n1, n2)
        X.copy_(R.reshape(k * n1, n2))
        Y = self.B._solve_rows_device(X)  # pylint: disable=protected-access  (rows (c, i1): B^{-1} X[c, i1, :])
        Yt = backend.alloc_matrix(k * n2, n1)
        Yt.copy_(Y.reshape(k, n1, n2).permute(0, 2, 1).reshape(k * n2, n1))
        Z = self.A._solve_rows_device(Yt)  # pylint: disable=protected-access  (rows (c, i2): A^{-1} Y[c, :, i2])
        return Z.reshape(k, n2, n1).permute(0, 2, 1).reshape(k, n1 * n2)

    def logabsdet(self) -> float:
        return self.B.shape[0] * self.A.logabsdet() + self.A.shape[0] * self.B.logabsdet()  # _kronecker.py:159-167

    def det(self) -> float:
        return float(self.A.det() ** self.B.shape[0] * self.B.det() ** self.A.shape[0])  # _kronecker.py:152-157

    def trace(self) -> float:
        return self.A.trace() * self.B.trace()

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
        self._blocks = [[aslinop(b) for b in row] for row in blocks]
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

    @property
    def T(self):
        return BlockMatrix([[self._blocks[i][j].T for i in range(len(self._blocks))] for j in range(len(self._blocks[0]))])

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


class ConcatenatedLinearOperator(LinearOperator):
    """Operators stacked along ``axis`` 0 (rows) or 1 (columns) (src/linpde_gp/linops/_concatenated.py:8-72)."""

    def __init__(self, linops, axis: int):
        linops = tuple(aslinop(op) for op in linops)
        if len(linops) < 1:
            raise ValueError("At least one linear operator must be given.")
        if axis not in [0, 1, -1, -2]:
            raise ValueError(f"axis is {axis}, expected one of 0, 1, -1, -2.")
        if axis < 0:
            axis += 2
        other = 1 - axis
        if not all(op.shape[other] == linops[0].shape[other] for op in linops):
            raise ValueError("All operators must agree along the axis that is not concatenated.")
        shape = [0, 0]
        shape[axis] = sum(op.shape[axis] for op in linops)
        shape[other] = linops[0].shape[other]
        super().__init__(shape)
        self._linops = linops
        self._axis = axis

    @property
    def linops(self):
        return self._linops

    @property
    def axis(self) -> int:
        return self._axis

    @property
    def T(self):
        return ConcatenatedLinearOperator(tuple(op.T for op in self._linops), 1 - self._axis)

    def device_dense(self):
        out = backend.alloc_matrix(*self.shape)
        o = 0
        for op in self._linops:
            n = op.shape[self._axis]
            op.assemble_into(out[o : o + n, :] if self._axis == 0 else out[:, o : o + n])
            o += n
        return out


class BlockMatrix2x2(LinearOperator):
    """``[[A, B], [C, D]]`` (src/linpde_gp/linops/_block.py:84-292) with the reference's four structures:

    * block diagonal (``B = C = None``);
    * symmetric positive definite (``is_spd=True``, exactly one of ``B``/``C`` given, the other one is its transpose):
      the bordered factor ``[[L_A, 0], [(L_A^{-1} B)^T, chol(S)]]``, ``S = D - (L_A^{-1}B)^T (L_A^{-1}B)`` (:191-242),
      is produced on the device by ``lpgp_chol_append`` -- A's cached factor is copied into a factor with one more
      segment, the new rows ``[C | D]`` are written behind it and TRSM + SYRK + POTRF run on those rows only;
    * block lower / upper triangular (triangular ``A``, ``D`` and ``B`` resp. ``C`` missing): block substitution
      (:251-266);
    * general (both ``B`` and ``C`` given, no structure): dense products only -- there is no LU on the device path.
    """

    def __init__(self, A, B, C, D, is_spd: bool = False):
        self._A, self._D = aslinop(A), aslinop(D)
        nA, mA = self._A.shape
        nD, mD = self._D.shape
        super().__init__((nA + nD, mA + mD))
        self.is_block_diagonal = False
        if B is None and C is None:
            self._B, self._C = Zero((nA, mD)), Zero((nD, mA))
            self.is_block_diagonal = True
            if self._A.is_symmetric and self._D.is_symmetric:
                self.is_symmetric = True
                if is_spd or (self._A.is_positive_definite and self._D.is_positive_definite):
                    self.is_positive_definite = True
            if self._A.is_lower_triangular and self._D.is_lower_triangular:
                self.is_lower_triangular = True
            if self._A.is_upper_triangular and self._D.is_upper_triangular:
                self.is_upper_triangular = True
        elif is_spd:
            if (B is None) == (C is None):
                raise ValueError("exactly one of B and C must be given for an SPD block matrix")
            if not (self._A.is_symmetric and self._D.is_symmetric):
                raise ValueError("A and D must be flagged symmetric")
            if self._A.is_positive_definite is False:
                raise ValueError("A must be positive definite")
            if C is None:
                self._B = aslinop(B)
                self._C = self._B.T
            else:
                self._C = aslinop(C)
                self._B = self._C.T
            self.is_symmetric = True
            self.is_positive_definite = True
        elif self._A.is_lower_triangular and self._D.is_lower_triangular and B is None:
            self._C = aslinop(C)
            self._B = Zero((nA, mD))
            self.is_lower_triangular = True
        elif self._A.is_upper_triangular and self._D.is_upper_triangular and C is None:
            self._B = aslinop(B)
            self._C = Zero((nD, mA))
            self.is_upper_triangular = True
        else:
            if B is None or C is None:
                raise ValueError("B and C must both be given unless the matrix is block diagonal, SPD or block triangular")
            self._B, self._C = aslinop(B), aslinop(C)
        if self._B.shape != (nA, mD) or self._C.shape != (nD, mA):
            raise ValueError("inconsistent block shapes")
        self._schur = None
        self._L_A_inv_B = None

    A = property(lambda self: self._A)
    B = property(lambda self: self._B)
    C = property(lambda self: self._C)
    D = property(lambda self: self._D)

    def _split(self, x: np.ndarray, axis: int = -2):
        return np.split(x, [self._A.shape[1]], axis=axis)

    @property
    def T(self):
        if self.is_symmetric:
            return self
        return BlockMatrix2x2(self._A.T, None if self.is_upper_triangular or self.is_block_diagonal else self._C.T,
                              None if self.is_lower_triangular or self.is_block_diagonal else self._B.T, self._D.T)

    def device_dense(self):
        out = backend.alloc_matrix(*self.shape)
        nA, mA = self._A.shape
        self._A.assemble_into(out[:nA, :mA])
        self._B.assemble_into(out[:nA, mA:])
        self._C.assemble_into(out[nA:, :mA])
        self._D.assemble_into(out[nA:, mA:])
        return out

    # -- SPD: bordered Cholesky on the device ------------------------------------------------------------------------
    def _require_spd(self):
        if not (self.is_symmetric and self.is_positive_definite):
            raise ValueError("This quantity can only be computed for SPD matrices.")

    def cholesky(self, lower: bool = True):
        if not (self.is_symmetric and self.is_positive_definite is not False):
            return super().cholesky(lower)
        if self._factor is None:
            fa = self._A.cholesky(True)  # cached on A: conditioning on a new batch reuses the old factor
            nA, nD = self._A.shape[0], self._D.shape[0]
            pA, pD = fa.factor.n, nD + nD % 2
            new = fa.factor.extended(pD)
            rows = new.L[pA : pA + pD]
            rows.zero_()
            Cd = self._C.device_dense()[:nD, :nA]
            rows[:nD].index_copy_(1, fa._dev_index(), Cd.contiguous())  # pylint: disable=protected-access
            self._D.assemble_into(rows[:nD, pA : pA + nD], lower=True)
            if nD % 2:
                rows[nD, pA + nD] = 1.0
            try:
                new.append_last()
            except np.linalg.LinAlgError:
                self.is_positive_definite = False
                raise
            self._factor = CholeskyFactor(new, index=np.concatenate([fa.index, pA + np.arange(nD)]))
        return self._factor if lower else self._factor.T

    @property
    def L_A_inv_B(self) -> LinearOperator:
        """``L_A^{-1} B`` (:203-207): the transposed off-diagonal block of the bordered factor."""
        self._require_spd()
        if self._L_A_inv_B is None:
            fac = self.cholesky(True)
            nA = self._A.shape[0]
            idx = fac._dev_index()  # pylint: disable=protected-access
            L21 = fac.factor.L[idx[nA:]][:, idx[:nA]]
            dev = backend.alloc_matrix(nA, self._D.shape[0])
            dev.copy_(L21.T)
            self._L_A_inv_B = _Device(dev)
        return self._L_A_inv_B

    @property
    def schur(self) -> LinearOperator:
        """Schur complement ``D - C A^{-1} B`` (:191-201)."""
        if self._schur is None:
            if self.is_symmetric and self.is_positive_definite:
                X = self.L_A_inv_B.device_dense()  # (nA, nD)
                Xt = backend.alloc_matrix(X.shape[1], X.shape[0])
                Xt.copy_(X.T)
                S = backend.alloc_matrix(*self._D.shape)
                self._D.assemble_into(S)
                backend.gemm_nt(Xt, Xt, S, -1.0, 1.0)
            else:
                AinvB = self._A.solve(self._B.todense())
                S = backend.to_device(self._D.todense() - self._C @ AinvB)
            self._schur = _Device(S)
            # the Schur complement of a symmetric (positive definite) matrix is symmetric (positive definite); in
            # every other case it is NOT flagged symmetric, so that solves with it are refused rather than wrong
            self._schur.is_symmetric = bool(self.is_symmetric)
            self._schur.is_positive_definite = True if (self.is_symmetric and self.is_positive_definite) else None
        return self._schur

    def schur_update(self, A_inv_u: np.ndarray, v: np.ndarray) -> np.ndarray:
        """Solution of ``[[A, B], [C, D]] [x; y] = [u; v]`` from ``A^{-1} u`` (:226-231)."""
        A_inv_u, v = np.asarray(A_inv_u, dtype=np.double), np.asarray(v, dtype=np.double)
        if self.is_block_diagonal:
            return np.concatenate((A_inv_u, self._D.solve(v)))
        y = self.schur.solve(v - self._C @ A_inv_u)
        x = A_inv_u - self._A.solve(self._B @ y)
        return np.concatenate((x, y))

    def solve(self, B):
        B = np.asarray(B, dtype=np.double)
        if B.shape[0 if B.ndim == 1 else -2] != self.shape[0]:
            raise ValueError("`b` must be a vector or a (stack of) matrices.")
        axis = 0 if B.ndim == 1 else -2
        b0, b1 = np.split(B, [self._A.shape[0]], axis=axis)
        if self.is_block_diagonal:
            return np.concatenate((self._A.solve(b0), self._D.solve(b1)), axis=axis)
        if self.is_symmetric:
            return self.cholesky(True).solve_spd(B)
        if self.is_lower_triangular:
            y0 = self._A.solve(b0)
            y1 = self._D.solve(b1 - self._C @ y0)
            return np.concatenate((y0, y1), axis=axis)
        if self.is_upper_triangular:
            y1 = self._D.solve(b1)
            y0 = self._A.solve(b0 - self._B @ y1)
            return np.concatenate((y0, y1), axis=axis)
        raise NotImplementedError("general (unstructured) block systems need an LU factorisation, which the device path "
                                  "does not provide")

    def _triangular_solve(self, B, trans: bool = False):
        return (self.T if trans else self).solve(B)

    def trace(self) -> float:
        return self._A.trace() + self._D.trace()

    def logabsdet(self) -> float:
        if self.is_block_diagonal or self.is_lower_triangular or self.is_upper_triangular:
            return self._A.logabsdet() + self._D.logabsdet()
        if self.is_symmetric and self.is_positive_definite:
            return self.cholesky(True).factor.logdet()
        raise NotImplementedError("determinant of an unstructured block matrix")

    def det(self) -> float:
        if self.is_block_diagonal or self.is_lower_triangular or self.is_upper_triangular:
            return self._A.det() * self._D.det()
        return float(np.exp(self.logabsdet()))


class _Device(LinearOperator):
    """A matrix that already lives on the device."""

    def __init__(self, dev: torch.Tensor):
        super().__init__(dev.shape)
        self._dev = dev

    def device_dense(self):
        return self._dev
