"""Linear functionals of the hot path (``linpde_gp.linfunctls``): point evaluation composed with an operator,
Lebesgue integrals, and their arithmetic.

``_EvaluationFunctional`` (src/linpde_gp/linfunctls/_evaluation.py:10-60) evaluates a function at the points ``X``
(output layout: codomain_shape + batch_shape); ``CompositeLinearFunctional`` (``_arithmetic.py:92-174``) is
``linop @ linfunctl @ linfuncop`` (matrix, functional, function operator; ``LinearFunctionOperator.to_linfunctl(X)``
returns one without a matrix); ``LebesgueIntegral``
(``_integrals.py:13-62``) integrates over an interval; ``ScaledLinearFunctional`` / ``SumLinearFunctional``
(``_arithmetic.py:12-90``) combine them, e.g. the stationarity condition of
experiments/0000_cpu_stationary_1d.ipynb cell 65.

For the conditioning code every functional flattens into a list of ATOMS ``(coef, kind, op, payload)`` with
``kind = "pts"`` (``coef * (op f)(X)``, payload = points) or ``"int"`` (``coef * int_a^b (op f)``, payload = (a, b));
all atoms of one functional produce the same number of rows."""
from __future__ import annotations

import numpy as np

from ..functions import Constant, Function, Polynomial, _as_shape


class LinearFunctional:
    def __init__(self, input_shapes, output_shape):
        self._input_domain_shape = _as_shape(input_shapes[0])
        self._input_codomain_shape = _as_shape(input_shapes[1])
        self._output_shape = _as_shape(output_shape)

    @property
    def input_shapes(self):
        return (self._input_domain_shape, self._input_codomain_shape)

    @property
    def input_domain_shape(self):
        return self._input_domain_shape

    @property
    def input_codomain_shape(self):
        return self._input_codomain_shape

    @property
    def output_shape(self):
        return self._output_shape

    @property
    def output_size(self):
        return int(np.prod(self._output_shape)) if self._output_shape else 1

    # (operator or None, evaluation points) -- what the conditioning code needs from a point-evaluation functional
    def _as_observation(self):
        atoms = self._atoms()
        if len(atoms) != 1 or atoms[0][1] != "pts" or atoms[0][0] != 1.0:
            raise NotImplementedError(f"{type(self).__name__} is not a plain point-evaluation observation")
        return atoms[0][2], atoms[0][3]

    def _atoms(self):  # pragma: no cover - abstract
        raise NotImplementedError(f"{type(self).__name__} cannot be used as an observation")

    def __call__(self, f, /, **kwargs):
        from ..randprocs import _conditional, _gaussian_process, covfuncs, crosscov

        if isinstance(f, covfuncs.CovarianceFunction):
            # L(k, argnum=1) = Cov(f(.), L[f]); argnum=0: the same object indexed the other way round (`reverse`)
            # (src/linpde_gp/randprocs/covfuncs/linfunctls/_registry.py, crosscov/linfunctls/_evaluation.py:21-328)
            argnum = kwargs.get("argnum", 0)
            if argnum not in (0, 1):
                raise ValueError("`argnum` must either be 0 or 1.")
            if isinstance(f, covfuncs.Zero):
                return crosscov.Zero(f.input_shape, f.output_shape_0, self.output_shape, reverse=(argnum == 0))
            return crosscov._FunctionalCrossCovariance(f, self, reverse=(argnum == 0))  # pylint: disable=protected-access
        if isinstance(f, crosscov.ProcessVectorCrossCovariance):
            return crosscov.apply_linfunctl(self, f)
        if isinstance(f, _conditional.ConditionalGaussianProcess):
            return f._apply_linfunctl(self)  # pylint: disable=protected-access
        if isinstance(f, _gaussian_process.GaussianProcess):
            return _gaussian_process.apply_linfunctl_to_gp(self, f)
        res = None
        for coef, kind, op, payload in self._atoms():
            g = f if op is None else op(f)
            v = g(payload) if kind == "pts" else _integrate_function(g, payload)
            v = v if coef == 1.0 else coef * np.asarray(v)
            res = v if res is None else res + v
        return res

    # -- arithmetic (src/linpde_gp/linfunctls/_linfunctl.py:74-129) ------------------------------------------------
    __array_ufunc__ = None

    def __neg__(self):
        return -1.0 * self

    def __add__(self, other):
        if isinstance(other, LinearFunctional):
            return SumLinearFunctional(self, other)
        return NotImplemented

    def __sub__(self, other):
        if isinstance(other, LinearFunctional):
            return self + (-other)
        return NotImplemented

    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledLinearFunctional(linfunctl=self, scalar=other)
        return NotImplemented

    def __matmul__(self, other):
        from ..linfuncops import LinearFunctionOperator

        if isinstance(other, LinearFunctionOperator):
            return CompositeLinearFunctional(linop=None, linfunctl=self, linfuncop=other)
        return NotImplemented

    def __rmatmul__(self, other):
        """``A @ linfunctl`` for a matrix ``A`` acting on the (1-D) output (``_linfunctl.py:117-129``)."""
        from .. import linops  # pylint: disable=import-outside-toplevel

        if isinstance(other, (np.ndarray, linops.LinearOperator)):
            return CompositeLinearFunctional(linop=other, linfunctl=self, linfuncop=None)
        return NotImplemented


def _integrate_function(g: Function, dom):
    """``LebesgueIntegral.__call__`` on functions (_integrals.py:36-62): closed form for constants, scipy.integrate.quad
    otherwise (the user's prior-mean function is host code; this is not on the device path)."""
    a, b = dom
    if isinstance(g, Constant):
        return g.value * (b - a)
    if isinstance(g, Polynomial):  # src/linpde_gp/functions/_linfunctls.py:9-13
        G = g.integrate()
        return G(np.asarray(b, dtype=np.double)) - G(np.asarray(a, dtype=np.double))
    import scipy.integrate  # pylint: disable=import-outside-toplevel

    if tuple(g.output_shape) != ():
        raise NotImplementedError("quadrature of vector-valued functions")
    return scipy.integrate.quad(lambda t: float(g(np.asarray(t, dtype=np.double))), a=a, b=b)[0]


class _EvaluationFunctional(LinearFunctional):
    def __init__(self, input_domain_shape, input_codomain_shape, X):
        from ..randprocs import covfuncs

        # an intact TensorProductGrid is remembered: conditioning assembles its Gram blocks from Kronecker factors
        self._grid = X if covfuncs._grid_factors(X) is not None else None  # pylint: disable=protected-access
        X = np.asarray(X, dtype=np.double)
        input_domain_shape = _as_shape(input_domain_shape)
        input_codomain_shape = _as_shape(input_codomain_shape)
        nd = len(input_domain_shape)
        if X.shape[X.ndim - nd :] != input_domain_shape:
            raise ValueError(f"trailing shape of X {X.shape} must equal the input domain shape {input_domain_shape}")
        batch = X.shape[: X.ndim - nd]
        super().__init__((input_domain_shape, input_codomain_shape), input_codomain_shape + batch)
        self._X = X

    @property
    def X(self):
        return self._X

    def _atoms(self):
        return [(1.0, "pts", None, self._X if self._grid is None else self._grid)]


class DiracFunctional(LinearFunctional):
    """Point evaluation with output layout ``batch_shape + codomain_shape`` (src/linpde_gp/linfunctls/_dirac.py:10-45;
    ``_EvaluationFunctional`` orders its output ``codomain_shape + batch_shape``).  For scalar-valued functions the
    two coincide; vector-valued processes are observed through ``_EvaluationFunctional`` / ``SelectOutput``."""

    def __init__(self, input_domain_shape, input_codomain_shape, X):
        X = np.asarray(X, dtype=np.double)
        input_domain_shape = _as_shape(input_domain_shape)
        input_codomain_shape = _as_shape(input_codomain_shape)
        nd = len(input_domain_shape)
        if X.shape[X.ndim - nd:] != input_domain_shape:
            raise ValueError(f"trailing shape of X {X.shape} must equal the input domain shape {input_domain_shape}")
        self._X = X
        self._X_batch_shape = X.shape[: X.ndim - nd]
        super().__init__((input_domain_shape, input_codomain_shape), self._X_batch_shape + input_codomain_shape)

    X = property(lambda self: self._X)
    X_batch_shape = property(lambda self: self._X_batch_shape)
    X_batch_ndim = property(lambda self: len(self._X_batch_shape))

    def _atoms(self):
        if self._input_codomain_shape != ():
            raise NotImplementedError("DiracFunctional observations of vector-valued processes")
        return [(1.0, "pts", None, self._X)]


class CompositeLinearFunctional(LinearFunctional):
    """``linop @ linfunctl @ linfuncop`` (``_arithmetic.py:92-174``): the function operator ``linfuncop`` acts first, then
    the functional, then the finite-dimensional matrix ``linop`` on the functional's (1-D) output -- the reference's
    keywords and meaning.  ``linfunctl @ L`` and ``A @ linfunctl`` build these (``_linfunctl.py:103-129``).

    Without a matrix the object flattens into the atoms of ``linfunctl`` with ``linfuncop`` composed in (the observation
    functionals of the conditioning path, ``LinearFunctionOperator.to_linfunctl``).  With a matrix it can be applied to
    functions, covariance functions (-> ``crosscov.LinOpProcessVectorCrossCovariance``), cross-covariances
    (-> ``Covariance``) and Gaussian processes (-> ``Normal``); conditioning a process on such a functional is not lowered
    to the device path (``NotImplementedError``; the reference's own use of it, the mass-matrix normaliser of L2
    projections, is built into the projection atoms)."""

    def __init__(self, *, linop=None, linfunctl, linfuncop=None):
        from ..linfuncops import LinearFunctionOperator

        if isinstance(linop, LinearFunctionOperator):  # round-1 spelling of this class: the operator passed as `linop`
            if linfuncop is not None:
                raise TypeError("`linop` is the matrix applied last; pass the function operator as `linfuncop`")
            linop, linfuncop = None, linop
        if linfuncop is not None and (
                tuple(linfuncop.output_shapes[0]) != tuple(linfunctl.input_domain_shape)
                or tuple(linfuncop.output_shapes[1]) != tuple(linfunctl.input_codomain_shape)):
            raise ValueError("shape mismatch between operator output and functional input")
        if linop is not None:
            from .. import linops  # pylint: disable=import-outside-toplevel

            linop = linop.todense() if isinstance(linop, linops.LinearOperator) else np.asarray(linop, dtype=np.double)
            if linop.ndim != 2 or len(linfunctl.output_shape) != 1 or linop.shape[1:] != tuple(linfunctl.output_shape):
                raise ValueError(f"a matrix of shape {linop.shape} cannot act on the output (shape "
                                 f"{tuple(linfunctl.output_shape)}) of the functional")
        super().__init__(linfunctl.input_shapes if linfuncop is None else linfuncop.input_shapes,
                         linfunctl.output_shape if linop is None else linop.shape[0:1])
        self._matrix = linop
        self._linfuncop = linfuncop
        self._linfunctl = linfunctl

    @property
    def linop(self):
        """The matrix applied last (an array; ``None`` if absent)."""
        return self._matrix

    @property
    def linfuncop(self):
        return self._linfuncop

    @property
    def linfunctl(self):
        return self._linfunctl

    def _without_matrix(self):
        if self._linfuncop is None:
            return self._linfunctl
        return CompositeLinearFunctional(linop=None, linfunctl=self._linfunctl, linfuncop=self._linfuncop)

    def _atoms(self):
        if self._matrix is not None:
            raise NotImplementedError("observations of the form matrix @ functional are not lowered to the device path; "
                                      "condition on the functional itself and transform the data instead")
        if self._linfuncop is None:
            return self._linfunctl._atoms()  # pylint: disable=protected-access
        # (inner.op @ linfuncop): the function operator of this composite acts first
        return [(c, kind, self._linfuncop if op is None else op @ self._linfuncop, payload)
                for c, kind, op, payload in self._linfunctl._atoms()]  # pylint: disable=protected-access

    def __call__(self, f, /, **kwargs):
        if self._matrix is None:
            return super().__call__(f, **kwargs)
        from .. import randvars  # pylint: disable=import-outside-toplevel
        from ..randprocs import crosscov  # pylint: disable=import-outside-toplevel

        A = self._matrix
        res = self._without_matrix()(f, **kwargs)
        if isinstance(res, crosscov.ProcessVectorCrossCovariance):  # covfuncs/linfunctls/_registry.py:42-61
            return crosscov.LinOpProcessVectorCrossCovariance(A, res)
        if isinstance(res, randvars.Covariance):  # f was a cross-covariance: A acts on this functional's axis
            M = np.asarray(res.matrix)
            if isinstance(f, crosscov.ProcessVectorCrossCovariance) and f.reverse:  # (randvar, functional) layout
                return randvars.ArrayCovariance((M @ A.T).reshape(res.shape0 + self.output_shape), res.shape0,
                                                self.output_shape)
            return randvars.ArrayCovariance((A @ M).reshape(self.output_shape + res.shape1), self.output_shape, res.shape1)
        if isinstance(res, randvars.Normal):
            return randvars.Normal(A @ res.mean.reshape(-1), A @ res.dense_cov @ A.T)
        return np.tensordot(np.asarray(res), A, axes=([-1], [1]))  # functions: A along the last axis (_arithmetic.py:150-151)

    def __matmul__(self, other):
        from ..linfuncops import LinearFunctionOperator

        if isinstance(other, LinearFunctionOperator):
            return CompositeLinearFunctional(
                linop=self._matrix, linfunctl=self._linfunctl,
                linfuncop=other if self._linfuncop is None else self._linfuncop @ other)
        return NotImplemented

    def __rmatmul__(self, other):
        from .. import linops  # pylint: disable=import-outside-toplevel

        if isinstance(other, (np.ndarray, linops.LinearOperator)):
            other = other.todense() if isinstance(other, linops.LinearOperator) else np.asarray(other, dtype=np.double)
            return CompositeLinearFunctional(linop=other if self._matrix is None else other @ self._matrix,
                                             linfunctl=self._linfunctl, linfuncop=self._linfuncop)
        return NotImplemented


class ScaledLinearFunctional(LinearFunctional):
    """``scalar * linfunctl`` (_arithmetic.py:12-55)."""

    def __init__(self, linfunctl, scalar):
        if np.ndim(scalar) != 0:
            raise ValueError()
        super().__init__(linfunctl.input_shapes, linfunctl.output_shape)
        self._linfunctl = linfunctl
        self._scalar = np.asarray(scalar, dtype=np.double)

    @property
    def linfunctl(self):
        return self._linfunctl

    @property
    def scalar(self):
        return self._scalar

    def _atoms(self):
        s = float(self._scalar)
        return [(s * c, kind, op, payload) for c, kind, op, payload in self._linfunctl._atoms()]  # pylint: disable=protected-access

    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledLinearFunctional(linfunctl=self._linfunctl, scalar=np.asarray(other) * self._scalar)
        return NotImplemented


class SumLinearFunctional(LinearFunctional):
    """``L_1 + ... + L_n`` (_arithmetic.py:58-90): all summands share input shapes and output shape."""

    def __init__(self, *summands):
        self._summands = tuple(summands)
        first = self._summands[0]
        if not all(tuple(s.input_domain_shape) == tuple(first.input_domain_shape)
                   and tuple(s.input_codomain_shape) == tuple(first.input_codomain_shape)
                   and tuple(s.output_shape) == tuple(first.output_shape) for s in self._summands):
            raise ValueError("summands of a SumLinearFunctional must agree in input and output shapes")
        super().__init__(first.input_shapes, first.output_shape)

    @property
    def summands(self):
        return self._summands

    def _atoms(self):
        return [a for s in self._summands for a in s._atoms()]  # pylint: disable=protected-access


class LebesgueIntegral(LinearFunctional):
    """``f -> int_domain f(x) dx`` over an interval (src/linpde_gp/linfunctls/_integrals.py:13-62).  Closed-form
    cross-covariances exist for univariate half-integer Matern kernels
    (src/linpde_gp/randprocs/covfuncs/linfunctls/_registry.py:175-193); there is no quadrature fallback on the device."""

    def __init__(self, input_domain, input_codomain_shape=()):
        dom = np.asarray(input_domain, dtype=np.double)
        if dom.shape != (2,):
            if dom.ndim == 2 and dom.shape[1] == 2:
                raise NotImplementedError("integrals over boxes have no closed form on the device (1-D intervals only)")
            raise TypeError("`input_domain` must be an interval (a, b)")
        if not dom[0] <= dom[1]:
            raise ValueError("empty interval")
        self._domain = (float(dom[0]), float(dom[1]))
        super().__init__(input_shapes=((), input_codomain_shape), output_shape=input_codomain_shape)

    @property
    def domain(self):
        return self._domain

    def _atoms(self):
        return [(1.0, "int", None, self._domain)]
