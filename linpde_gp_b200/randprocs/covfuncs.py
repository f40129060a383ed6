"""Covariance functions of the hot path with the reference's API; numerics on the GPU.

Mirrors ``linpde_gp.randprocs.covfuncs`` / ``probnum.randprocs.covfuncs``:
``CovarianceFunction.__call__/matrix/linop`` (pn/randprocs/covfuncs/_covariance_function.py:280-488),
``Matern`` (pn …/_matern.py), ``ExpQuad`` (pn …/_exponentiated_quadratic.py), ``TensorProduct``
(src/linpde_gp/randprocs/covfuncs/_tensor_product.py:15-95), the scalar / sum wrappers
(src/linpde_gp/randprocs/covfuncs/_jax_arithmetic.py:16-66) and the operator-transformed kernels selected by
the dispatch registry src/linpde_gp/randprocs/covfuncs/linfuncops/diffops/_registry.py.

Every kernel that can be evaluated is *lowered* to a flat descriptor (``_lowering.py``) and evaluated by the
hand-written CUDA kernels behind ``liblpgp.so``; there is no numpy evaluation path.
"""
from __future__ import annotations

import numpy as np

from .. import _lowering, backend
from ..functions import _as_shape
from ..linops import CovarianceLinearOperator, Kronecker


class TensorProductGrid(np.ndarray):
    """Points of a tensor-product grid, ``grid[i_1, ..., i_d] = (f_1[i_1], ..., f_d[i_d])``, that remember their 1-D
    factors (src/linpde_gp/randprocs/covfuncs/_tensor_product.py:133-152).  Tensor-product kernels evaluated on two
    such grids give Kronecker-structured matrices; ``linop`` exploits that (see :class:`TensorProduct`)."""

    def __new__(cls, *factors, indexing="ij"):
        factors = tuple(np.asarray(f, dtype=np.double) for f in factors)
        if not factors or any(f.ndim != 1 for f in factors):
            raise ValueError("the factors of a TensorProductGrid must be one-dimensional arrays")
        mesh = np.meshgrid(*factors, copy=True, sparse=False, indexing=indexing)
        obj = np.stack(mesh, axis=-1).view(cls)
        obj.factors = factors
        obj.indexing = indexing
        return obj

    def __array_finalize__(self, obj):
        # slices / arithmetic results are plain point sets: they do not inherit the factorisation
        self.factors = None
        self.indexing = None


def _grid_factors(x):
    """The 1-D factors of ``x`` if it is an intact "ij"-ordered TensorProductGrid (C-order flattening of the grid
    equals the Kronecker ordering of the factors), else ``None``."""
    if isinstance(x, TensorProductGrid) and getattr(x, "factors", None) is not None and x.indexing == "ij":
        return x.factors
    return None


def _kron_all(ops):
    out = ops[0]
    for o in ops[1:]:
        out = Kronecker(out, o)
    return out


class CovarianceFunction:
    """Scalar-output covariance function ``k: R^input_shape x R^input_shape -> R``."""

    def __init__(self, input_shape=()):
        self._input_shape = _as_shape(input_shape)

    # -- shapes ------------------------------------------------------------------------------------------
    @property
    def input_shape(self):
        return self._input_shape

    input_shape_0 = input_shape
    input_shape_1 = input_shape

    @property
    def input_ndim(self):
        return len(self._input_shape)

    @property
    def input_size(self):
        return int(np.prod(self._input_shape)) if self._input_shape else 1

    # scalar-output unless a subclass (multi-output kernels below) overrides these
    @property
    def output_shape_0(self):
        return ()

    @property
    def output_shape_1(self):
        return ()

    @property
    def output_shape(self):
        return self.output_shape_0

    @property
    def output_size_0(self):
        return int(np.prod(self.output_shape_0)) if self.output_shape_0 else 1

    @property
    def output_size_1(self):
        return int(np.prod(self.output_shape_1)) if self.output_shape_1 else 1

    # -- lowering ------------------------------------------------------------------------------------------
    def _product_form(self):
        """``(factors, L0_terms, L1_terms, scale)`` or raise ``NotImplementedError``."""
        raise NotImplementedError(f"{type(self).__name__} has no closed product form on the device path")

    def _radial_form(self):
        """``dict(nu, scales, dir0, dir1, sigma2)`` for kernels built on an isotropic multi-dimensional Matern kernel
        (not of product form; evaluated by the radial device family), else ``None``."""
        return None

    def descriptor(self):
        if getattr(self, "_desc", None) is None:
            radial = self._radial_form()
            if radial is not None:
                self._desc = _lowering.lower_radial(**radial)
            else:
                factors, t0, t1, scale = self._product_form()
                self._desc = _lowering.lower(factors, t0, t1, scale)
        return self._desc

    # -- evaluation ----------------------------------------------------------------------------------------
    def _check_shapes(self, x0_shape, x1_shape=None):
        err = (
            "The shape of the input array `x{argnum}` must match `input_shape_{argnum}`, i.e. `{input_shape}`, "
            "along its trailing dimensions, but an array with shape `{shape}` was given."
        )
        nd = self.input_ndim
        if tuple(x0_shape[len(x0_shape) - nd :]) != self._input_shape:
            raise ValueError(err.format(argnum=0, input_shape=self._input_shape, shape=x0_shape))
        batch = tuple(x0_shape[: len(x0_shape) - nd])
        if x1_shape is not None:
            if tuple(x1_shape[len(x1_shape) - nd :]) != self._input_shape:
                raise ValueError(err.format(argnum=1, input_shape=self._input_shape, shape=x1_shape))
            try:
                batch = np.broadcast_shapes(batch, tuple(x1_shape[: len(x1_shape) - nd]))
            except ValueError as ve:
                raise ValueError(
                    f"The input arrays `x0` and `x1` with shapes {x0_shape} and {x1_shape} can not be broadcast "
                    "to a common shape."
                ) from ve
        return batch

    def __call__(self, x0, x1=None):
        """``k(x0, x1)`` with numpy broadcasting over batch shapes; ``x1=None`` evaluates ``k(x0_i, x0_i)``
        (pn …/_covariance_function.py:280-357)."""
        x0 = np.asarray(x0, dtype=np.double)
        x1 = None if x1 is None else np.asarray(x1, dtype=np.double)
        batch = self._check_shapes(x0.shape, None if x1 is None else x1.shape)
        return self._evaluate(x0, x1, batch)

    def _evaluate(self, x0, x1, batch):
        desc = self.descriptor()
        d = self.input_size
        nd = self.input_ndim
        if x1 is None:
            n = int(np.prod(batch)) if batch else 1
            return backend.gram_diag(desc, n).cpu().numpy().reshape(batch)
        b0, b1 = x0.shape[: x0.ndim - nd], x1.shape[: x1.ndim - nd]
        nb = len(batch)
        b0 = (1,) * (nb - len(b0)) + tuple(b0)
        b1 = (1,) * (nb - len(b1)) + tuple(b1)
        # outer-product pattern: every batch axis varies in at most one of the arguments
        if all(s0 == 1 or s1 == 1 for s0, s1 in zip(b0, b1)):
            X0 = backend.points(x0, d)
            X1 = backend.points(x1, d)
            K = backend.gram(desc, X0, X1).cpu().numpy()  # (prod b0, prod b1)
            # pair up axis i of x0's batch with axis i of x1's batch (one of them is a singleton) and merge
            K = K.reshape(tuple(b0) + tuple(b1))
            perm = [ax for pair in zip(range(nb), range(nb, 2 * nb)) for ax in pair]
            return np.ascontiguousarray(np.transpose(K, perm)).reshape(batch)
        full0 = np.broadcast_to(x0, batch + self._input_shape).reshape(-1, d)
        full1 = np.broadcast_to(x1, batch + self._input_shape).reshape(-1, d)
        out = backend.gram_pairs(desc, backend.to_device(full0), backend.to_device(full1))
        return out.cpu().numpy().reshape(batch)

    def _preprocess_linop_input(self, x, argnum):
        x = np.asarray(x, dtype=np.double)
        nd = self.input_ndim
        if not (x.ndim >= nd and x.shape[x.ndim - nd :] == self._input_shape):
            raise ValueError(
                f"The shape of `x{argnum}` must must match `input_shape_{argnum}`, i.e. `{self._input_shape}`, of "
                f"the covariance function along its trailing dimensions, but an array with shape `{x.shape}` was given."
            )
        return x.reshape((-1,) + self._input_shape, order="C")

    def linop(self, x0, x1=None) -> CovarianceLinearOperator:
        """Lazy device-resident covariance matrix (pn …/_covariance_function.py:415-488); ``x1=None`` means
        ``x1 := x0`` (full symmetric matrix), unlike ``__call__``."""
        x0 = self._preprocess_linop_input(x0, 0)
        x1 = None if x1 is None else self._preprocess_linop_input(x1, 1)
        return CovarianceLinearOperator(self, x0, x1)

    def matrix(self, x0, x1=None) -> np.ndarray:
        """Dense covariance matrix as a host array (pn …/_covariance_function.py:359-413)."""
        return self.linop(x0, x1).todense()

    # -- arithmetic ----------------------------------------------------------------------------------------
    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledCovarianceFunction(self, scalar=other)
        return NotImplemented

    def __add__(self, other):
        if isinstance(other, CovarianceFunction):
            return SumCovarianceFunction(self, other)
        return NotImplemented


class _Stationary1DMixin:
    def _factors(self):
        raise NotImplementedError


class Matern(CovarianceFunction):
    """Matern covariance function (pn/randprocs/covfuncs/_matern.py:25-195); half-integer ``nu`` on the device."""

    def __init__(self, input_shape=(), nu=1.5, *, lengthscales=None):
        super().__init__(input_shape)
        nu = float(nu)
        if nu <= 0:
            raise ValueError(f"Hyperparameter nu={nu} must be positive.")
        self._nu = nu
        ls = np.asarray(1.0 if lengthscales is None else lengthscales, dtype=np.double)
        if np.any(ls <= 0):
            raise ValueError(f"All lengthscales l={ls} must be positive.")
        np.broadcast_to(ls, self.input_shape)
        self._lengthscales = ls

    @property
    def nu(self):
        return self._nu

    @property
    def p(self):
        q = self._nu - 0.5
        return int(q) if q == int(q) else None

    @property
    def is_half_integer(self):
        return self.p is not None

    @property
    def lengthscales(self):
        return self._lengthscales

    def _radial_form(self):
        if self.input_size == 1:
            return None
        if self.input_ndim != 1:
            raise NotImplementedError("Matern inputs must be scalars or vectors")
        if not self.is_half_integer:
            raise NotImplementedError("only half-integer Matern kernels are supported on the device")
        ls = np.broadcast_to(self._lengthscales, (self.input_size,))
        return {"nu": self._nu, "scales": np.sqrt(2.0 * self._nu) / ls, "dir0": None, "dir1": None, "sigma2": 1.0}

    def _product_form(self):
        if self.input_size != 1:
            raise NotImplementedError(
                "isotropic multi-dimensional Matern kernels are not of product form (they lower to the radial "
                "device family: plain evaluation and first-order directional derivatives, like the reference)"
            )
        if not self.is_half_integer:
            raise NotImplementedError("only half-integer Matern kernels are supported on the device")
        ell = float(np.broadcast_to(self._lengthscales, (1,))[0])
        return [_lowering.Factor1D("matern", ell, nu=self._nu)], None, None, 1.0


class ExpQuad(CovarianceFunction):
    """Exponentiated quadratic (pn/randprocs/covfuncs/_exponentiated_quadratic.py:14-100), ARD lengthscales."""

    def __init__(self, input_shape=(), *, lengthscales=None):
        super().__init__(input_shape)
        ls = np.asarray(1.0 if lengthscales is None else lengthscales, dtype=np.double)
        if np.any(ls <= 0):
            raise ValueError(f"Lengthscales l={ls} must be positive.")
        np.broadcast_to(ls, self.input_shape)
        self._lengthscales = ls

    @property
    def lengthscales(self):
        return self._lengthscales

    def _product_form(self):
        if self.input_ndim > 1:
            raise NotImplementedError("ExpQuad inputs must be scalars or vectors")
        ls = np.broadcast_to(self._lengthscales, (self.input_size,))
        return [_lowering.Factor1D("expquad", float(l)) for l in ls], None, None, 1.0


class TensorProduct(CovarianceFunction):
    """Product of univariate kernels (src/linpde_gp/randprocs/covfuncs/_tensor_product.py:15-48)."""

    def __init__(self, *factors):
        if len(factors) < 1:
            raise ValueError("At least one factor is required.")
        if not all(isinstance(k, CovarianceFunction) and k.input_shape == () for k in factors):
            raise ValueError("The input shape of all factors must be `()`.")
        self._factors = tuple(factors)
        super().__init__((len(self._factors),))

    @property
    def factors(self):
        return self._factors

    def linop(self, x0, x1=None):
        """Kronecker product of the factors' 1-D covariance matrices when both inputs are tensor-product grids
        (src/linpde_gp/randprocs/covfuncs/_tensor_product.py:64-82), else the generic lazy covariance matrix."""
        f0, f1 = _grid_factors(x0), (None if x1 is None else _grid_factors(x1))
        if f0 is not None and len(f0) == len(self._factors) and (x1 is None or (f1 is not None and len(f1) == len(f0))):
            return _kron_all([k.linop(f0[i], None if x1 is None else f1[i]) for i, k in enumerate(self._factors)])
        return super().linop(x0, x1)

    def _product_form(self):
        fs, scale = [], 1.0
        for k in self._factors:
            f, t0, t1, s = k._product_form()
            if t0 is not None or t1 is not None or len(f) != 1:
                raise NotImplementedError("TensorProduct factors must be plain univariate kernels")
            fs.extend(f)
            scale *= s
        return fs, None, None, scale


class ScaledCovarianceFunction(CovarianceFunction):
    """``scalar * k`` (reference: ``JaxScaledCovarianceFunction``, _jax_arithmetic.py:16-44)."""

    def __init__(self, covfunc, scalar):
        if np.ndim(scalar) != 0:
            raise TypeError()
        super().__init__(covfunc.input_shape)
        self._covfunc = covfunc
        self._scalar = np.asarray(scalar, dtype=np.double)

    @property
    def scalar(self):
        return self._scalar

    @property
    def covfunc(self):
        return self._covfunc

    @property
    def output_shape_0(self):
        return self._covfunc.output_shape_0

    @property
    def output_shape_1(self):
        return self._covfunc.output_shape_1

    def _product_form(self):
        f, t0, t1, s = self._covfunc._product_form()
        return f, t0, t1, s * float(self._scalar)

    def _radial_form(self):
        radial = self._covfunc._radial_form()
        if radial is not None:
            radial = dict(radial, sigma2=radial["sigma2"] * float(self._scalar))
        return radial

    def _evaluate(self, x0, x1, batch):
        if self.output_shape_0 != () or self.output_shape_1 != ():
            return float(self._scalar) * self._covfunc._evaluate(x0, x1, batch)
        return super()._evaluate(x0, x1, batch)

    def linop(self, x0, x1=None):
        if _grid_factors(x0) is not None:  # keep the Kronecker structure of the wrapped kernel
            inner = self._covfunc.linop(x0, x1)
            if inner.kron_terms() is not None:
                return float(self._scalar) * inner
        return super().linop(x0, x1)

    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return ScaledCovarianceFunction(self._covfunc, scalar=np.asarray(other) * self._scalar)
        return NotImplemented


# reference-compatible aliases (src/linpde_gp/randprocs/covfuncs/__init__.py:1-19)
JaxScaledCovarianceFunction = ScaledCovarianceFunction


class SumCovarianceFunction(CovarianceFunction):
    """``k1 + k2 + ...`` (reference: ``JaxSumCovarianceFunction``, _jax_arithmetic.py:47-66).  Summands sharing
    the same base factors are folded into ONE device descriptor; otherwise blocks are accumulated."""

    def __init__(self, *summands):
        if not all(isinstance(s, CovarianceFunction) for s in summands):
            raise TypeError()
        if not all(s.input_shape == summands[0].input_shape for s in summands):
            raise ValueError("All summands must have the same input shape")
        if not all(s.output_shape_0 == summands[0].output_shape_0 and s.output_shape_1 == summands[0].output_shape_1
                   for s in summands):
            raise ValueError("All summands must have the same output shapes")
        super().__init__(summands[0].input_shape)
        self._summands = tuple(summands)

    @property
    def summands(self):
        return self._summands

    @property
    def output_shape_0(self):
        return self._summands[0].output_shape_0

    @property
    def output_shape_1(self):
        return self._summands[0].output_shape_1

    def _product_form(self):
        raise NotImplementedError("sum of kernels with different base factors: evaluated block-wise")

    def descriptors(self):
        return device_descriptors(self)

    def linop(self, x0, x1=None):
        if _grid_factors(x0) is not None:
            parts = [s.linop(x0, x1) for s in self._summands]
            if all(p.kron_terms() is not None for p in parts):
                out = parts[0]
                for p in parts[1:]:
                    out = out + p
                return out
        return super().linop(x0, x1)

    def _evaluate(self, x0, x1, batch):
        out = None
        for s in self._summands:
            v = s._evaluate(x0, x1, batch)
            out = v if out is None else out + v
        return out


JaxSumCovarianceFunction = SumCovarianceFunction


class LinDiffOpCovarianceFunction(CovarianceFunction):
    """``L0 k L1^*`` for a product-form base kernel ``k`` and partial-derivative operators ``L0`` / ``L1``.

    Generic carrier of what the reference spreads over ``TensorProduct_LinDiffOp_LinDiffOp`` and the
    ``ExpQuad_* / *HalfIntegerMatern_*`` closed-form classes (diffops/_tensor_product.py, _expquad.py, _matern.py);
    the named subclasses below only record WHICH reference class the dispatcher would have produced."""

    def __init__(self, k, *, L0=None, L1=None):
        super().__init__(k.input_shape)
        self._k = k
        self._L0 = L0
        self._L1 = L1

    @property
    def k(self):
        return self._k

    @property
    def L0(self):
        return self._L0

    @property
    def L1(self):
        return self._L1

    def _radial_form(self):
        radial = self._k._radial_form()
        if radial is None:
            return None
        d = self._k.input_size

        def direction(L):
            """First-order operator -> its direction vector (the reference dispatches DirectionalDerivative only,
            diffops/_registry.py:142-190; anything else on an isotropic multi-d Matern kernel needs its jax fallback)."""
            if L is None:
                return None
            vec = np.zeros(d)
            for mi, c in L._terms().items():
                if len(mi) != d:
                    raise ValueError("operator and kernel input dimensions differ")
                if sum(mi) != 1:
                    raise NotImplementedError(
                        "isotropic multi-dimensional Matern kernels have closed forms for first-order (directional) "
                        "derivatives only (diffops/_registry.py:270-280: the Laplacian falls to the reference's jax path)")
                vec[mi.index(1)] += c
            return vec

        return dict(radial, dir0=direction(self._L0), dir1=direction(self._L1))

    def _product_form(self):
        f, t0, t1, s = self._k._product_form()
        assert t0 is None and t1 is None
        d = len(f)

        def terms(L):
            if L is None:
                return None
            t = L._terms()
            if any(len(mi) != d for mi in t):
                raise ValueError("operator and kernel input dimensions differ")
            return t

        return f, terms(self._L0), terms(self._L1), s


class _UnivariateDerivativeFactor(CovarianceFunction):
    """``d^a/dx^a d^b/dx'^b kappa(x, x')`` of ONE univariate factor of a tensor-product kernel: what the reference's
    ``k_x0_x1s`` cache holds (diffops/_tensor_product.py:34-82).  Lowered to a 1-D device descriptor."""

    def __init__(self, factor: CovarianceFunction, a: int, b: int):
        super().__init__(())
        self._factor, self._a, self._b = factor, int(a), int(b)

    def _product_form(self):
        f, t0, t1, s = self._factor._product_form()
        if t0 is not None or t1 is not None or len(f) != 1:
            raise NotImplementedError("TensorProduct factors must be plain univariate kernels")
        return f, ({(self._a,): 1.0} if self._a else None), ({(self._b,): 1.0} if self._b else None), s


class TensorProduct_LinDiffOp_LinDiffOp(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_tensor_product.py:21"""

    def linop(self, x0, x1=None):
        """Sum over the operator terms of Kronecker products of 1-D derivative-kernel matrices when the inputs are
        tensor-product grids (diffops/_tensor_product.py:84-119, 140-156), else the generic lazy matrix.

        Deviation from the reference, on purpose: with ``x1`` given the reference pairs ``x0``'s factors with
        THEMSELVES (``x1_factors = x0.factors``, diffops/_tensor_product.py:147-148, an upstream slip); here the
        factors of ``x1`` are used, so that ``linop(x0, x1).todense() == matrix(x0, x1)`` always holds."""
        f0, f1 = _grid_factors(x0), (None if x1 is None else _grid_factors(x1))
        base = self._k
        d = len(base.factors) if isinstance(base, TensorProduct) else 0
        if d and f0 is not None and len(f0) == d and (x1 is None or (f1 is not None and len(f1) == d)):
            ident = {(0,) * d: 1.0}
            t0 = ident if self._L0 is None else self._L0._terms()
            t1 = ident if self._L1 is None else self._L1._terms()
            cache = {}

            def factor_linop(i, a, b):
                if (i, a, b) not in cache:
                    kf = _UnivariateDerivativeFactor(base.factors[i], a, b)
                    cache[(i, a, b)] = kf.linop(f0[i], None if x1 is None else f1[i])
                return cache[(i, a, b)]

            res = 0
            for mi0, c0 in t0.items():
                for mi1, c1 in t1.items():
                    term = _kron_all([factor_linop(i, mi0[i], mi1[i]) for i in range(d)])
                    res = res + (float(c0) * float(c1)) * term
            return res
        return super().linop(x0, x1)


class ExpQuad_Identity_DirectionalDerivative(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_expquad.py:12"""


class ExpQuad_DirectionalDerivative_DirectionalDerivative(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_expquad.py:76"""


class ExpQuad_Identity_WeightedLaplacian(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_expquad.py:145"""


class ExpQuad_WeightedLaplacian_WeightedLaplacian(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_expquad.py:221"""


class ExpQuad_DirectionalDerivative_WeightedLaplacian(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_expquad.py:348"""


class HalfIntegerMatern_Identity_DirectionalDerivative(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_matern.py:17"""


class HalfIntegerMatern_DirectionalDerivative_DirectionalDerivative(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_matern.py:138 (input_size > 1: radial device family)"""


class UnivariateHalfIntegerMatern_DirectionalDerivative_DirectionalDerivative(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_matern.py:267"""


class UnivariateHalfIntegerMatern_Identity_WeightedLaplacian(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_matern.py:358"""


class UnivariateHalfIntegerMatern_WeightedLaplacian_WeightedLaplacian(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_matern.py:439"""


class UnivariateHalfIntegerMatern_DirectionalDerivative_WeightedLaplacian(LinDiffOpCovarianceFunction):  # pylint: disable=invalid-name
    """diffops/_matern.py:512"""



# ---------------------------------------------------------------------------------------------------------------
# multi-output kernels
# ---------------------------------------------------------------------------------------------------------------
class Zero(CovarianceFunction):
    """The zero covariance function (src/linpde_gp/randprocs/covfuncs/_zero.py): the cross-covariance between
    independent outputs.  Contributes no device work -- blocks made of it are zero-filled."""

    def __init__(self, input_shape=()):
        super().__init__(input_shape)

    def _product_form(self):
        raise NotImplementedError("the zero kernel has no descriptor (nothing to evaluate)")

    def _evaluate(self, x0, x1, batch):
        return np.zeros(batch)

    def linop(self, x0, x1=None):
        from ..linops import Matrix

        n0 = self._preprocess_linop_input(x0, 0).shape[0]
        n1 = n0 if x1 is None else self._preprocess_linop_input(x1, 1).shape[0]
        return Matrix(np.zeros((n0, n1)))

    def __rmul__(self, other):
        return self if np.ndim(other) == 0 else NotImplemented


class IndependentMultiOutputCovarianceFunction(CovarianceFunction):
    """``k(x, x')[i, j] = delta_ij k_i(x, x')`` for scalar kernels ``k_i`` on a common input space
    (src/linpde_gp/randprocs/covfuncs/_independent_multi_output.py:11-71)."""

    def __init__(self, *covfuncs):
        if not covfuncs:
            raise ValueError("at least one covariance function is needed")
        for c in covfuncs:
            if not isinstance(c, CovarianceFunction):
                raise TypeError()
            if c.input_shape != covfuncs[0].input_shape or c.output_shape_0 != () or c.output_shape_1 != ():
                raise ValueError("the outputs must be scalar kernels on the same input space")
        super().__init__(covfuncs[0].input_shape)
        self._covfuncs = tuple(covfuncs)

    @property
    def covfuncs(self):
        return self._covfuncs

    @property
    def output_shape_0(self):
        return (len(self._covfuncs),)

    @property
    def output_shape_1(self):
        return (len(self._covfuncs),)

    def _product_form(self):
        raise NotImplementedError("multi-output kernel: select an output first")

    def _evaluate(self, x0, x1, batch):
        n = len(self._covfuncs)
        out = np.zeros(tuple(batch) + (n, n))
        for i, c in enumerate(self._covfuncs):
            out[..., i, i] = c._evaluate(x0, x1, batch)
        return out

    def linop(self, x0, x1=None):
        from ..linops import BlockDiagonalMatrix

        return BlockDiagonalMatrix(*(c.linop(x0, x1) for c in self._covfuncs))


class StackCovarianceFunction(CovarianceFunction):
    """A vector of scalar kernels read as ONE kernel with a vector-valued output on argument ``output_idx``
    (src/linpde_gp/randprocs/covfuncs/_stack.py:15-104): what ``L(k_multi, argnum=1)`` is for a scalar-valued
    observation operator ``L`` -- the covariance between every output of the process and the observed quantity."""

    def __init__(self, covfuncs, output_idx: int = 1):
        covfuncs = tuple(covfuncs)
        if not covfuncs or not all(isinstance(c, CovarianceFunction) for c in covfuncs):
            raise ValueError()
        if any(c.input_shape != covfuncs[0].input_shape for c in covfuncs):
            raise ValueError()
        if any(c.output_shape_0 != () or c.output_shape_1 != () for c in covfuncs):
            raise ValueError()
        output_idx = int(output_idx)
        if output_idx not in (0, 1):
            raise ValueError()
        super().__init__(covfuncs[0].input_shape)
        self._covfuncs = covfuncs
        self._output_idx = output_idx

    @property
    def covfuncs(self):
        return self._covfuncs

    @property
    def output_idx(self):
        return self._output_idx

    @property
    def output_shape_0(self):
        return (len(self._covfuncs),) if self._output_idx == 0 else ()

    @property
    def output_shape_1(self):
        return (len(self._covfuncs),) if self._output_idx == 1 else ()

    def _product_form(self):
        raise NotImplementedError("multi-output kernel: select an output first")

    def _evaluate(self, x0, x1, batch):
        return np.stack([c._evaluate(x0, x1, batch) for c in self._covfuncs], axis=-1)

    def linop(self, x0, x1=None):
        from ..linops import BlockMatrix

        ops = [c.linop(x0, x1) for c in self._covfuncs]
        return BlockMatrix([[o] for o in ops] if self._output_idx == 0 else [ops])

    # sums / scalar multiples of stacks stay stacks (element-wise), so that one SelectOutput resolves them
    def __add__(self, other):
        if isinstance(other, StackCovarianceFunction):
            if other.output_idx != self._output_idx or len(other.covfuncs) != len(self._covfuncs):
                raise ValueError("stacks of different layout cannot be added")
            return StackCovarianceFunction(
                tuple(_add_cov(a, b) for a, b in zip(self._covfuncs, other.covfuncs)), output_idx=self._output_idx)
        return super().__add__(other)

    def __rmul__(self, other):
        if np.ndim(other) == 0:
            return StackCovarianceFunction(tuple(_scale_cov(other, c) for c in self._covfuncs), output_idx=self._output_idx)
        return NotImplemented


def _add_cov(a: CovarianceFunction, b: CovarianceFunction) -> CovarianceFunction:
    """``a + b`` with zeros dropped and nested sums flattened."""
    if isinstance(a, Zero):
        return b
    if isinstance(b, Zero):
        return a
    if isinstance(a, StackCovarianceFunction) or isinstance(b, StackCovarianceFunction):
        return a + b
    parts = []
    for c in (a, b):
        parts.extend(c.summands if isinstance(c, SumCovarianceFunction) else (c,))
    return SumCovarianceFunction(*parts)


def _scale_cov(scalar, k: CovarianceFunction) -> CovarianceFunction:
    """``scalar * k`` distributed over sums (every summand keeps its own device descriptor)."""
    if isinstance(k, (Zero, StackCovarianceFunction)):
        return scalar * k
    if isinstance(k, SumCovarianceFunction):
        return SumCovarianceFunction(*(_scale_cov(scalar, s) for s in k.summands))
    return scalar * k


def _apply_multi_output(L, k: CovarianceFunction, argnum: int) -> CovarianceFunction:
    """Operators and kernels with vector-valued (co)domains: linearity over sums / scalars / compositions, then the
    ``SelectOutput`` and element-wise ``StackCovarianceFunction`` rules of
    src/linpde_gp/randprocs/covfuncs/linfuncops/_registry.py:34-48, 82-120."""
    from ..linfuncops import _linfuncop

    if isinstance(L, _linfuncop.CompositeLinearFunctionOperator):
        out = k
        for op in reversed(L.linfuncops):
            out = apply_linfuncop(op, out, argnum)
        return out
    if isinstance(L, _linfuncop.SumLinearFunctionOperator) and L.input_codomain_shape != ():
        out = None
        for summand in L.summands:
            v = apply_linfuncop(summand, k, argnum)
            out = v if out is None else _add_cov(out, v)
        return out
    if isinstance(L, _linfuncop.ScaledLinearFunctionOperator) and L.input_codomain_shape != ():
        return _scale_cov(float(L.scalar), apply_linfuncop(L.linfuncop, k, argnum))
    if isinstance(k, ScaledCovarianceFunction):
        return _scale_cov(float(k.scalar), apply_linfuncop(L, k.covfunc, argnum))
    if isinstance(k, SumCovarianceFunction):
        out = None
        for s in k.summands:
            v = apply_linfuncop(L, s, argnum)
            out = v if out is None else _add_cov(out, v)
        return out
    if isinstance(L, _linfuncop.SelectOutput):
        if isinstance(k, IndependentMultiOutputCovarianceFunction):
            if not isinstance(L.idx, (int, np.integer)):
                raise NotImplementedError("SelectOutput with a non-integer index")
            zero = Zero(k.input_shape)
            return StackCovarianceFunction(
                tuple(c if i == L.idx else zero for i, c in enumerate(k.covfuncs)), output_idx=1 - argnum)
        if isinstance(k, StackCovarianceFunction) and k.output_idx == argnum:
            return k.covfuncs[L.idx]
        raise NotImplementedError(f"SelectOutput on argument {argnum} of {type(k).__name__}")
    if isinstance(k, StackCovarianceFunction):
        if (argnum == 0 and k.output_idx == 1) or (argnum == 1 and k.output_idx == 0):
            return StackCovarianceFunction(tuple(apply_linfuncop(L, c, argnum) for c in k.covfuncs), output_idx=k.output_idx)
        raise NotImplementedError("the operator acts on the stacked argument: select an output first")
    raise NotImplementedError(f"{type(L).__name__} applied to {type(k).__name__}")


def device_descriptors(k: CovarianceFunction):
    """Device descriptors whose values add up to ``k``: one for a kernel with a product form, else one per summand,
    with outer scalars distributed over sums and ``Zero`` summands dropped."""
    if isinstance(k, Zero):
        return []
    if isinstance(k, SumCovarianceFunction):
        out = []
        for s in k.summands:
            out.extend(device_descriptors(s))
        return out
    if isinstance(k, ScaledCovarianceFunction) and isinstance(k.covfunc, (SumCovarianceFunction, Zero, ScaledCovarianceFunction)):
        inner = k.covfunc
        if isinstance(inner, ScaledCovarianceFunction):
            return device_descriptors(ScaledCovarianceFunction(inner.covfunc, scalar=float(k.scalar) * float(inner.scalar)))
        return device_descriptors(_scale_cov(float(k.scalar), inner))
    return [k.descriptor()]


def _kind_of(L):
    from ..linfuncops import diffops

    if L is None:
        return "Identity"
    if isinstance(L, diffops.ScaledLinearDifferentialOperator):
        return _kind_of(L.lindiffop)
    if isinstance(L, diffops.WeightedLaplacian):
        return "WeightedLaplacian"
    if isinstance(L, diffops.DirectionalDerivative):
        return "DirectionalDerivative"
    if isinstance(L, diffops.PartialDerivative) and L.order == 0:
        return "Identity"
    return "LinDiffOp"


def _select_class(k, L0, L1):
    """Which closed-form class the reference's registry produces for (L0, k, L1)."""
    if isinstance(k, TensorProduct):
        return TensorProduct_LinDiffOp_LinDiffOp
    kinds = tuple(sorted((_kind_of(L0), _kind_of(L1))))
    table_eq = {
        ("DirectionalDerivative", "Identity"): ExpQuad_Identity_DirectionalDerivative,
        ("DirectionalDerivative", "DirectionalDerivative"): ExpQuad_DirectionalDerivative_DirectionalDerivative,
        ("Identity", "WeightedLaplacian"): ExpQuad_Identity_WeightedLaplacian,
        ("WeightedLaplacian", "WeightedLaplacian"): ExpQuad_WeightedLaplacian_WeightedLaplacian,
        ("DirectionalDerivative", "WeightedLaplacian"): ExpQuad_DirectionalDerivative_WeightedLaplacian,
    }
    table_m = {
        ("DirectionalDerivative", "Identity"): HalfIntegerMatern_Identity_DirectionalDerivative,
        ("DirectionalDerivative", "DirectionalDerivative"): UnivariateHalfIntegerMatern_DirectionalDerivative_DirectionalDerivative,
        ("Identity", "WeightedLaplacian"): UnivariateHalfIntegerMatern_Identity_WeightedLaplacian,
        ("WeightedLaplacian", "WeightedLaplacian"): UnivariateHalfIntegerMatern_WeightedLaplacian_WeightedLaplacian,
        ("DirectionalDerivative", "WeightedLaplacian"): UnivariateHalfIntegerMatern_DirectionalDerivative_WeightedLaplacian,
    }
    if isinstance(k, ExpQuad):
        return table_eq.get(kinds, LinDiffOpCovarianceFunction)
    if isinstance(k, Matern):
        if k.input_size > 1 and kinds == ("DirectionalDerivative", "DirectionalDerivative"):
            return HalfIntegerMatern_DirectionalDerivative_DirectionalDerivative  # diffops/_registry.py:156-190
        return table_m.get(kinds, LinDiffOpCovarianceFunction)
    return LinDiffOpCovarianceFunction


def _compose(L_old, L_new):
    """Operators acting on the same argument compose; only ``identity`` is composable without a product rule."""
    if L_old is None:
        return L_new
    raise NotImplementedError(
        "an operator was already applied to this argument (the reference returns NotImplemented as well, "
        "diffops/_registry.py:54-72)"
    )


def apply_linfuncop(L, k: CovarianceFunction, argnum: int = 0) -> CovarianceFunction:
    """``L(k, argnum=...)``: the dispatch of src/linpde_gp/randprocs/covfuncs/linfuncops/_registry.py:14-31 (scalar
    and sum linearity) and diffops/_registry.py:15-370 (closed forms)."""
    from ..linfuncops import _linfuncop

    if argnum not in (0, 1):
        raise ValueError("`argnum` must either be 0 or 1.")
    if tuple(L.input_domain_shape) != tuple(k.input_shape):
        raise ValueError(f"operator input domain shape {L.input_domain_shape} != kernel input shape {k.input_shape}")
    k_out = k.output_shape_0 if argnum == 0 else k.output_shape_1
    if tuple(L.input_codomain_shape) != tuple(k_out):
        raise ValueError(f"operator input codomain shape {L.input_codomain_shape} != kernel output shape {k_out}")
    if isinstance(k, Zero):
        return k
    if (L.input_codomain_shape != () or k.output_shape_0 != () or k.output_shape_1 != ()
            or isinstance(L, _linfuncop.CompositeLinearFunctionOperator)):
        return _apply_multi_output(L, k, argnum)
    if isinstance(L, _linfuncop.Identity):
        return k
    if isinstance(k, ScaledCovarianceFunction):
        return k.scalar * apply_linfuncop(L, k.covfunc, argnum)
    if isinstance(k, SumCovarianceFunction):
        return SumCovarianceFunction(*(apply_linfuncop(L, s, argnum) for s in k.summands))
    if isinstance(k, LinDiffOpCovarianceFunction):
        L0, L1 = k.L0, k.L1
        if argnum == 0:
            L0 = _compose(L0, L)
        else:
            L1 = _compose(L1, L)
        out = _select_class(k.k, L0, L1)(k.k, L0=L0, L1=L1)
        out._radial_form()  # radial kernels: first-order operators only
        return out
    if isinstance(k, (Matern, ExpQuad, TensorProduct)):
        L0, L1 = (L, None) if argnum == 0 else (None, L)
        out = _select_class(k, L0, L1)(k, L0=L0, L1=L1)
        # validate now: raise NotImplementedError for unsupported combinations
        if out._radial_form() is None:
            out._product_form()
        return out
    raise NotImplementedError(f"{type(L).__name__} applied to {type(k).__name__}")
