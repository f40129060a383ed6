"""Process-vector cross-covariances ``x -> Cov(f(x), L[f])`` (``linpde_gp.randprocs.crosscov``).

The reference represents the N-column object ``(k L*)(., X)`` that posterior evaluation and the off-diagonal Gram
blocks are built from as a ``ProcessVectorCrossCovariance`` (src/linpde_gp/randprocs/crosscov/_pv_crosscov.py:14-161):
``linfunctl(k, argnum=1)`` creates one (``covfuncs/linfunctls/_registry.py``, ``crosscov/linfunctls/_evaluation.py``),
``pv(x)`` evaluates it as an array of shape ``batch + randvar_shape`` (``reverse=False``) or ``randvar_shape + batch``
(``reverse=True``), ``pv.evaluate_linop(x)`` as an ``(M, N)`` / ``(N, M)`` linear operator, and a second functional
applied to it, ``linfunctl(pv)``, yields the ``Covariance`` between the two random vectors
(``crosscov/linfunctls/_evaluation.py:11-18``).

Here every such object flattens into ATOMS ``(coef, kind, op, payload)`` of its functional (see
``linpde_gp_b200/linfunctls.py``); evaluation runs on the device: point atoms through the pairwise Gram kernel
(``lpgp_gram``), integral atoms through the closed-form Matern integral kernel (``lpgp_matern_integral``).  Scalar-output
processes only (multi-output cross-covariances are handled inside ``ConditionalGaussianProcess``)."""
from __future__ import annotations

import numpy as np

from .. import backend, linops, randvars
from ..functions import _as_shape


class ProcessVectorCrossCovariance:
    """Base class mirroring ``_pv_crosscov.py:14-161``: shapes, ``__call__`` / ``evaluate_linop`` with shape checks,
    scalar multiples and sums."""

    def __init__(self, randproc_input_shape, randproc_output_shape, randvar_shape, reverse: bool = True):
        self._randproc_input_shape = _as_shape(randproc_input_shape)
        self._randproc_output_shape = _as_shape(randproc_output_shape)
        self._randvar_shape = _as_shape(randvar_shape)
        self._reverse = bool(reverse)

    randproc_input_shape = property(lambda self: self._randproc_input_shape)
    randproc_input_ndim = property(lambda self: len(self._randproc_input_shape))
    randproc_output_shape = property(lambda self: self._randproc_output_shape)
    randproc_output_ndim = property(lambda self: len(self._randproc_output_shape))
    randvar_shape = property(lambda self: self._randvar_shape)
    randvar_ndim = property(lambda self: len(self._randvar_shape))
    randvar_size = property(lambda self: int(np.prod(self._randvar_shape)) if self._randvar_shape else 1)
    reverse = property(lambda self: self._reverse)

    def _check_input(self, x: np.ndarray):
        nd = self.randproc_input_ndim
        if x.shape[x.ndim - nd:] != self._randproc_input_shape:
            raise ValueError(
                "The shape of the input array must match the `randproc_input_shape` "
                f"`{self._randproc_input_shape}` of the function along its last dimensions, but an array with shape "
                f"`{x.shape}` was given."
            )
        return x.shape[: x.ndim - nd]

    # -- device evaluation: (M, N) matrix  Cov(f(x_i), (L f)_j)  for M flattened test points -----------------
    def _device_matrix(self, Xt, out=None, accumulate: bool = False, alpha: float = 1.0):  # pragma: no cover - abstract
        raise NotImplementedError

    def _points(self, x: np.ndarray):
        d = int(np.prod(self._randproc_input_shape)) if self._randproc_input_shape else 1
        return backend.points(np.asarray(x, dtype=np.double), d)

    def __call__(self, x) -> np.ndarray:
        x = np.asarray(x, dtype=np.double)
        batch = self._check_input(x)
        if self._randproc_output_shape != ():
            raise NotImplementedError("multi-output cross-covariances")
        K = self._device_matrix(self._points(x))[:, : self.randvar_size].cpu().numpy()  # (M, N)
        if self._reverse:
            return K.T.reshape(self._randvar_shape + batch)
        return K.reshape(batch + self._randvar_shape)

    def evaluate_linop(self, x) -> linops.LinearOperator:
        x = np.asarray(x, dtype=np.double)
        self._check_input(x)
        if self._randproc_output_shape != ():
            raise NotImplementedError("multi-output cross-covariances")
        K = self._device_matrix(self._points(x))
        op = linops._Device(K[:, : self.randvar_size])  # pylint: disable=protected-access
        return op.T if self._reverse else op

    # -- arithmetic (_pv_crosscov.py:163-194, _arithmetic.py) --------------------------------------------------
    __array_ufunc__ = None

    def __neg__(self):
        return -1.0 * self

    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledProcessVectorCrossCovariance(self, other)
        return NotImplemented

    def __add__(self, other):
        if isinstance(other, ProcessVectorCrossCovariance):
            return SumProcessVectorCrossCovariance(self, other)
        return NotImplemented

    def __sub__(self, other):
        if isinstance(other, ProcessVectorCrossCovariance):
            return self + (-other)
        return NotImplemented

    # -- a function operator / functional applied to the FREE argument -----------------------------------------
    def _apply_linfuncop(self, L):  # pragma: no cover - abstract
        raise NotImplementedError(f"{type(L).__name__} cannot be applied to {type(self).__name__}")

    def _atom_pairs(self):
        """``[(coef, covfunc, atom), ...]``: this object equals ``sum coef * Cov(f(.), atom[f])`` under ``covfunc``
        (operators already applied to the free argument are folded into ``covfunc``)."""
        raise NotImplementedError


class _FunctionalCrossCovariance(ProcessVectorCrossCovariance):
    """``Cov(f(.), linfunctl[f])`` for ``f ~ GP(., covfunc)``: the single implementation behind the reference's
    ``CovarianceFunction_Identity_Evaluation`` / ``..._Evaluation_Identity`` (crosscov/linfunctls/_evaluation.py:21-328)
    and ``CovarianceFunction_Identity_LebesgueIntegral`` (crosscov/linfunctls/integrals/)."""

    def __init__(self, covfunc, linfunctl, reverse: bool, *, free_op=None):
        from . import _conditional  # pylint: disable=import-outside-toplevel

        if covfunc.output_shape_0 != () or covfunc.output_shape_1 != ():
            raise NotImplementedError("multi-output cross-covariances")
        super().__init__(covfunc.input_shape, (), linfunctl.output_shape, reverse=reverse)
        self._covfunc = covfunc
        self._linfunctl = linfunctl
        self._free_op = free_op
        d = covfunc.input_size
        self._atoms = [_conditional._Atom(c, kind, op, payload, d)  # pylint: disable=protected-access
                       for c, kind, op, payload in linfunctl._atoms()]  # pylint: disable=protected-access
        if any(a.n != self.randvar_size for a in self._atoms):
            raise ValueError("all summands of the functional must produce `randvar_size` values")

    covfunc = property(lambda self: self._covfunc)
    linfunctl = property(lambda self: self._linfunctl)

    def _kernel_for(self, atom):
        """Kernel ``(free_op k atom.op*)`` with the test point as argument 0."""
        kk = atom.apply(self._covfunc, 1)
        return kk if self._free_op is None else self._free_op(kk, argnum=0)

    def _device_matrix(self, Xt, out=None, accumulate: bool = False, alpha: float = 1.0):
        from .. import _lowering  # pylint: disable=import-outside-toplevel
        from . import _conditional  # pylint: disable=import-outside-toplevel

        m, n = Xt.shape[0], self.randvar_size
        if out is None:
            out = backend.alloc_matrix(m, n)
            accumulate = False
        for atom in self._atoms:
            kk = self._kernel_for(atom)
            coef = alpha * atom.coef
            if atom.kind == "proj":
                blk = _conditional._proj_pts_block(self._covfunc, atom.proj, Xt, coef, op_pts=self._free_op, op_proj=atom.op)  # pylint: disable=protected-access
                if accumulate:
                    out[:, :n].add_(blk)
                else:
                    out[:, :n].copy_(blk)
            elif atom.kind == "pts":
                _conditional._gram_into(kk, Xt, atom.X, out[:, :n], accumulate=accumulate, alpha=coef)  # pylint: disable=protected-access
            else:
                terms = _conditional._integral_terms(kk)  # pylint: disable=protected-access
                if not terms and not accumulate:
                    out[:, :1].zero_()
                for t, (scale, nu, ell) in enumerate(terms):
                    dsc = _lowering.matern_integral_desc(nu, ell)
                    backend.matern_integral(dsc, atom.dom[0], atom.dom[1], Xt, out, out_stride=out.stride(0),
                                            alpha=coef * scale, accumulate=accumulate or t > 0)
            accumulate = True
        return out

    def _apply_linfuncop(self, L):
        op = L if self._free_op is None else L @ self._free_op
        return _FunctionalCrossCovariance(self._covfunc, self._linfunctl, self._reverse, free_op=op)

    def _atom_pairs(self):
        return [(1.0, self, None)]


class CovarianceFunction_L2Projection_UnivariateLinearInterpolationBasis(_FunctionalCrossCovariance):  # pylint: disable=invalid-name
    """``Cov(f(.), P[f])`` for an L2 projection ``P`` onto hat functions (crosscov/linfunctls/projections.py:18-69): what
    ``proj(k, argnum)`` returns.  Evaluation: closed-form hat integrals for half-integer Matern kernels of every order,
    Gauss-Legendre per element for smooth kernels (``_conditional._proj_pts_block``), normaliser included."""

    def __init__(self, covfunc, proj, reverse: bool = True):
        super().__init__(covfunc, proj, reverse)

    projection = property(lambda self: self._linfunctl)


class Matern32_L2Projection_UnivariateLinearInterpolationBasis(CovarianceFunction_L2Projection_UnivariateLinearInterpolationBasis):  # pylint: disable=invalid-name
    """The class the reference returns for ``nu = 3/2`` (projections.py:125-170, its only closed form); here the same
    device kernel serves every half-integer order."""


class ScaledProcessVectorCrossCovariance(ProcessVectorCrossCovariance):
    """``scalar * pv_crosscov`` (crosscov/_arithmetic.py:12-60)."""

    def __init__(self, pv_crosscov: ProcessVectorCrossCovariance, scalar):
        if np.ndim(scalar) != 0:
            raise ValueError("`scalar` must be a scalar")
        super().__init__(pv_crosscov.randproc_input_shape, pv_crosscov.randproc_output_shape, pv_crosscov.randvar_shape,
                         reverse=pv_crosscov.reverse)
        self._pv_crosscov = pv_crosscov
        self._scalar = float(scalar)

    pv_crosscov = property(lambda self: self._pv_crosscov)
    scalar = property(lambda self: self._scalar)

    def _device_matrix(self, Xt, out=None, accumulate: bool = False, alpha: float = 1.0):
        return self._pv_crosscov._device_matrix(Xt, out=out, accumulate=accumulate, alpha=alpha * self._scalar)  # pylint: disable=protected-access

    def _apply_linfuncop(self, L):
        return ScaledProcessVectorCrossCovariance(self._pv_crosscov._apply_linfuncop(L), self._scalar)  # pylint: disable=protected-access

    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledProcessVectorCrossCovariance(self._pv_crosscov, float(other) * self._scalar)
        return NotImplemented


class LinOpProcessVectorCrossCovariance(ProcessVectorCrossCovariance):
    """``linop @ pv_crosscov``: the cross-covariance with the random vector ``A L[f]`` for a matrix ``A`` acting on the
    (1-D) random vector of ``pv_crosscov`` (crosscov/_arithmetic.py:91-130) -- what ``(A @ linfunctl)(k, argnum)`` returns
    (covfuncs/linfunctls/_registry.py:42-61).  Evaluation: the inner (M x n) matrix on the device, then one DMMA GEMM
    with ``A`` (m x n)."""

    def __init__(self, linop, pv_crosscov: ProcessVectorCrossCovariance):
        A = linop.todense() if isinstance(linop, linops.LinearOperator) else np.asarray(linop, dtype=np.double)
        if pv_crosscov.randvar_ndim != 1 or A.ndim != 2 or A.shape[1:] != pv_crosscov.randvar_shape:
            raise ValueError(f"a matrix of shape {A.shape} cannot act on a random vector of shape {pv_crosscov.randvar_shape}")
        super().__init__(pv_crosscov.randproc_input_shape, pv_crosscov.randproc_output_shape, A.shape[0:1],
                         reverse=pv_crosscov.reverse)
        self._linop = A
        self._pv_crosscov = pv_crosscov
        self._A_dev = None

    linop = property(lambda self: self._linop)
    pv_crosscov = property(lambda self: self._pv_crosscov)

    def _device_matrix(self, Xt, out=None, accumulate: bool = False, alpha: float = 1.0):
        if self._A_dev is None:
            self._A_dev = backend.alloc_matrix(*self._linop.shape)
            self._A_dev.copy_(backend.to_device(self._linop))
        inner = self._pv_crosscov._device_matrix(Xt)[:, : self._pv_crosscov.randvar_size]  # pylint: disable=protected-access
        if out is None:
            out = backend.alloc_matrix(Xt.shape[0], self.randvar_size)
            accumulate = False
        backend.gemm_nt(inner, self._A_dev, out[:, : self.randvar_size], alpha=alpha, beta=1.0 if accumulate else 0.0)
        return out

    def _apply_linfuncop(self, L):
        return LinOpProcessVectorCrossCovariance(self._linop, self._pv_crosscov._apply_linfuncop(L))  # pylint: disable=protected-access


class SumProcessVectorCrossCovariance(ProcessVectorCrossCovariance):
    """``pv_1 + ... + pv_n`` (crosscov/_arithmetic.py:63-120); nested sums are flattened."""

    def __init__(self, *pv_crosscovs: ProcessVectorCrossCovariance):
        flat = []
        for pv in pv_crosscovs:
            flat.extend(pv.summands if isinstance(pv, SumProcessVectorCrossCovariance) else [pv])
        first = flat[0]
        if not all(pv.randproc_input_shape == first.randproc_input_shape
                   and pv.randproc_output_shape == first.randproc_output_shape
                   and pv.randvar_shape == first.randvar_shape and pv.reverse == first.reverse for pv in flat):
            raise ValueError("All summands must agree in shapes and in `reverse`.")
        super().__init__(first.randproc_input_shape, first.randproc_output_shape, first.randvar_shape, reverse=first.reverse)
        self._summands = tuple(flat)

    summands = property(lambda self: self._summands)

    def _device_matrix(self, Xt, out=None, accumulate: bool = False, alpha: float = 1.0):
        for i, pv in enumerate(self._summands):
            out = pv._device_matrix(Xt, out=out, accumulate=accumulate or i > 0, alpha=alpha)  # pylint: disable=protected-access
        return out

    def _apply_linfuncop(self, L):
        return SumProcessVectorCrossCovariance(*(pv._apply_linfuncop(L) for pv in self._summands))  # pylint: disable=protected-access


class Zero(ProcessVectorCrossCovariance):
    """The cross-covariance that vanishes identically (crosscov/_zero.py)."""

    def _device_matrix(self, Xt, out=None, accumulate: bool = False, alpha: float = 1.0):
        if out is None:
            out = backend.alloc_matrix(Xt.shape[0], self.randvar_size)
            accumulate = False
        return out if accumulate else out.zero_()

    def _apply_linfuncop(self, L):
        return self


class StackedProcessVectorCrossCovariance(ProcessVectorCrossCovariance):
    """Concatenation of cross-covariances along the random-vector axis: the reference's
    ``ConditionalGaussianProcess.PriorPredictiveCrossCovariance`` (_conditional.py:112-175), i.e. ``k(., [X_1 ... X_n])``
    over all observation batches.  ``reverse=False`` only."""

    def __init__(self, pv_crosscovs):
        pvs = tuple(pv_crosscovs)
        if not pvs:
            raise ValueError("at least one cross-covariance")
        first = pvs[0]
        if not all(pv.randproc_input_shape == first.randproc_input_shape and pv.randproc_output_shape == first.randproc_output_shape
                   and not pv.reverse for pv in pvs):
            raise ValueError("all cross-covariances must share the process shapes and have reverse=False")
        super().__init__(first.randproc_input_shape, first.randproc_output_shape, (sum(pv.randvar_size for pv in pvs),),
                         reverse=False)
        self._pvs = pvs

    pv_crosscovs = property(lambda self: self._pvs)

    def append(self, pv: ProcessVectorCrossCovariance) -> "StackedProcessVectorCrossCovariance":
        return StackedProcessVectorCrossCovariance(self._pvs + (pv,))

    def _device_matrix(self, Xt, out=None, accumulate: bool = False, alpha: float = 1.0):
        if out is None:
            out = backend.alloc_matrix(Xt.shape[0], self.randvar_size)
            accumulate = False
        off = 0
        for pv in self._pvs:
            n = pv.randvar_size
            if off % 2 == 0:
                pv._device_matrix(Xt, out=out[:, off : off + n], accumulate=accumulate, alpha=alpha)  # pylint: disable=protected-access
            else:  # the Gram kernel stores 16-byte aligned rows: odd offsets go through a temporary
                tmp = pv._device_matrix(Xt, alpha=alpha)  # pylint: disable=protected-access
                if accumulate:
                    out[:, off : off + n].add_(tmp[:, :n])
                else:
                    out[:, off : off + n].copy_(tmp[:, :n])
            off += n
        return out

    def _apply_linfuncop(self, L):
        return StackedProcessVectorCrossCovariance(tuple(pv._apply_linfuncop(L) for pv in self._pvs))  # pylint: disable=protected-access


# -- functional applied to the remaining argument: Cov(L0[f], L1[f]) ----------------------------------------------------
def apply_linfunctl(linfunctl, pv: ProcessVectorCrossCovariance) -> randvars.Covariance:
    """``linfunctl(pv_crosscov)`` (crosscov/linfunctls/_evaluation.py:11-18, _linfunctl.py:21-82): the covariance between
    ``linfunctl[f]`` and the random vector of ``pv_crosscov`` as a lazy ``LinearOperatorCovariance`` whose matrix is
    assembled on the device."""
    from . import _conditional  # pylint: disable=import-outside-toplevel

    if pv.randproc_output_shape != ():
        raise NotImplementedError("multi-output cross-covariances")
    d = int(np.prod(pv.randproc_input_shape)) if pv.randproc_input_shape else 1
    atoms = [_conditional._Atom(c, kind, op, payload, d) for c, kind, op, payload in linfunctl._atoms()]  # pylint: disable=protected-access
    n0, n1 = atoms[0].n, pv.randvar_size
    if any(a.n != n0 for a in atoms):
        raise ValueError("all summands of the functional must produce the same number of values")
    out = backend.alloc_matrix(n0, n1)
    first = True
    for atom in atoms:
        target = pv if atom.op is None else pv._apply_linfuncop(atom.op)  # pylint: disable=protected-access
        if atom.kind == "pts":
            target._device_matrix(atom.X, out=out, accumulate=not first, alpha=atom.coef)  # pylint: disable=protected-access
        else:  # integral / L2 projection of the free argument
            _integrate_free_argument(target, atom, out, accumulate=not first)
        first = False
    op = linops._Device(out[:, :n1])  # pylint: disable=protected-access
    shape0, shape1 = tuple(linfunctl.output_shape), tuple(pv.randvar_shape)
    if pv.reverse:
        return randvars.LinearOperatorCovariance(op.T, shape0=shape1, shape1=shape0)
    return randvars.LinearOperatorCovariance(op, shape0=shape0, shape1=shape1)


def _integrate_free_argument(pv, atom, out, accumulate: bool) -> None:
    """Row ``out[0, :] (+)= coef * int_a^b pv(t) dt`` for an integral atom applied to the free argument (closed forms of
    the (double) Matern integrals, crosscov/linfunctls/integrals/), or the ``m`` rows ``coef * P[pv]`` of an L2-projection
    atom (crosscov/linfunctls/projections.py:69-122), evaluated per atom of ``pv``."""
    from . import _conditional  # pylint: disable=import-outside-toplevel

    def walk(p, alpha):
        if isinstance(p, ScaledProcessVectorCrossCovariance):
            yield from walk(p.pv_crosscov, alpha * p.scalar)
        elif isinstance(p, SumProcessVectorCrossCovariance):
            for s in p.summands:
                yield from walk(s, alpha)
        elif isinstance(p, _FunctionalCrossCovariance):
            yield p, alpha
        elif isinstance(p, Zero):
            return
        else:
            raise NotImplementedError(f"integral of {type(p).__name__}")

    for p, alpha in walk(pv, 1.0):
        for b_atom in p._atoms:  # pylint: disable=protected-access
            a_atom = _conditional._Atom.__new__(_conditional._Atom)  # pylint: disable=protected-access
            a_atom.coef, a_atom.kind, a_atom.op, a_atom.X_host, a_atom.X, a_atom.n, a_atom.dom, a_atom.proj = (
                alpha * atom.coef, atom.kind, p._free_op, None, None, atom.n, atom.dom, atom.proj)  # pylint: disable=protected-access
            _conditional._atom_cov_into(p.covfunc, a_atom, b_atom, out[: atom.n, : b_atom.n], accumulate=accumulate)  # pylint: disable=protected-access
            accumulate = True
    if not accumulate:
        out[: atom.n].zero_()
