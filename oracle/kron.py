"""CPU restatement of the reference's tensor-grid (Kronecker) structure path.  TEST INFRASTRUCTURE ONLY: imported by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg, never by the product.

Follows, operation by operation:
  * ``TensorProductGrid``  (src/linpde_gp/randprocs/covfuncs/_tensor_product.py:133-152): meshgrid of the factors,
    ``indexing="ij"``, stacked on the last axis;
  * ``TensorProduct.linop`` on grids (_tensor_product.py:64-82): ``functools.reduce(pn.linops.Kronecker, [k_d.linop(
    x0.factors[d], x1.factors[d])])``;
  * ``TensorProduct_LinDiffOp_LinDiffOp.linop`` on grids
    (src/linpde_gp/randprocs/covfuncs/linfuncops/diffops/_tensor_product.py:84-119, 140-156): sum over the
    coefficient terms of L0 and L1 of ``c_a c_b`` times the Kronecker product of the 1-D derivative-kernel matrices;
  * ``pn.linops.Kronecker`` (pn/linops/_kronecker.py:17-166): ``todense = np.kron(A, B)`` and the vec-trick matvec
    ``(A (x) B) vec(X) = vec(A X B^T)`` for C-ordered (row-major) vec.

Parity pinned by ``tests/golden/kron.npz`` (``oracle/make_golden.py``, outputs of the real reference).  One deliberate
difference: the reference's derivative ``linop(x0, x1)`` pairs ``x0``'s factors with themselves when ``x1`` is given
(diffops/_tensor_product.py:147-148); like the product, the restatement uses ``x1``'s factors, so the golden vectors
of derivative kernels are taken with ``x1=None`` or from ``.matrix(x0, x1)``.
"""
from __future__ import annotations

import functools

import numpy as np

from . import covfuncs as ocf


def tensor_product_grid(*factors) -> np.ndarray:
    factors = [np.asarray(f, dtype=np.double) for f in factors]
    return np.stack(np.meshgrid(*factors, copy=True, sparse=False, indexing="ij"), axis=-1)


def _factor_matrix(f, a, b, x0, x1):
    x0 = np.asarray(x0, dtype=np.double)
    x1 = x0 if x1 is None else np.asarray(x1, dtype=np.double)
    return np.asarray(ocf.univariate_factor(f, a, b, x0[:, None], x1[None, :]))


def kronecker_terms(kernel, L0, L1, factors0, factors1=None):
    """``[(coefficient, [K_1, ..., K_d]), ...]`` with ``L0 k L1* (grid0, grid1) = sum coeff * K_1 (x) ... (x) K_d``."""
    base = kernel["base"]
    assert base["kind"] == "tensor_product"
    d = len(base["factors"])
    ident = {tuple([0] * d): 1.0}

    def coeffs(L):
        if L is None:
            return ident
        out = {}
        for c, op in L:
            for mi, v in ocf.op_coefficients(op).items():
                out[mi] = out.get(mi, 0.0) + c * v
        return out

    c0, c1 = coeffs(L0), coeffs(L1)
    scale = 1.0 if kernel.get("scale") is None else float(kernel["scale"])
    memo = {}
    terms = []
    for mi0, v0 in c0.items():
        for mi1, v1 in c1.items():
            mats = []
            for i in range(d):
                key = (i, mi0[i], mi1[i])
                if key not in memo:
                    memo[key] = _factor_matrix(base["factors"][i], mi0[i], mi1[i], factors0[i],
                                               None if factors1 is None else factors1[i])
                mats.append(memo[key])
            terms.append((scale * v0 * v1, mats))
    return terms


def dense(terms) -> np.ndarray:
    return sum(c * functools.reduce(np.kron, mats) for c, mats in terms)


def matvec(terms, x: np.ndarray) -> np.ndarray:
    """Structured product: every Kronecker term applied factor by factor (never densified)."""
    x = np.asarray(x, dtype=np.double)
    vec = x.ndim == 1
    X = x[:, None] if vec else x
    out = 0.0
    for c, mats in terms:
        shape_in = [m.shape[1] for m in mats]
        T = X.reshape(shape_in + [X.shape[1]])
        for ax, m in enumerate(mats):
            T = np.moveaxis(np.tensordot(m, T, axes=([1], [ax])), 0, ax)
        out = out + c * T.reshape(-1, X.shape[1])
    return out[:, 0] if vec else out
