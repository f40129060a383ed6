"""Oracle: dense FP64 factorisation and solves (numpy/scipy restatement).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference's arithmetic on this part of the path
is LAPACK ``dpotrf/dpotrs/dtrtrs`` reached through scipy (unpinned beyond ``scipy>=1.4``,
/root/reference/pyproject.toml:31-33); call sites: pn/linops/_linear_operator.py:296-315, 860-865 and
src/linpde_gp/linops/_block.py:191-268.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg


def cholesky_lower(A: np.ndarray) -> np.ndarray:
    """``LinearOperator._cholesky`` (pn/linops/_linear_operator.py:860-865): raises ``np.linalg.LinAlgError``
    if ``A`` is not positive definite."""
    return scipy.linalg.cholesky(A, lower=True)


def cho_solve_lower(L: np.ndarray, B: np.ndarray) -> np.ndarray:
    """``LinearOperator._solve`` SPD branch (pn/linops/_linear_operator.py:303-307)."""
    return scipy.linalg.cho_solve((L, True), B)


def solve_lower(L: np.ndarray, B: np.ndarray, trans: bool = False) -> np.ndarray:
    """Triangular branch (pn/linops/_linear_operator.py:296-299)."""
    return scipy.linalg.solve_triangular(L, B, lower=True, trans=1 if trans else 0)


def cholesky_append(L_A: np.ndarray, B: np.ndarray, D: np.ndarray) -> np.ndarray:
    """Bordered block Cholesky of [[A, B], [B^T, D]] given ``L_A = chol(A)``:
    ``L_A_inv_B`` (TRSM, _block.py:203-207), Schur complement ``D - (L_A^-1 B)^T (L_A^-1 B)`` (:191-201),
    its Cholesky and the assembled block factor (:233-242)."""
    n0, n1 = L_A.shape[0], D.shape[0]
    L_A_inv_B = solve_lower(L_A, B)
    S = D - L_A_inv_B.T @ L_A_inv_B
    L_S = cholesky_lower(S)
    L = np.zeros((n0 + n1, n0 + n1), dtype=np.double)
    L[:n0, :n0] = L_A
    L[n0:, :n0] = L_A_inv_B.T
    L[n0:, n0:] = L_S
    return L


def schur_update(L_A: np.ndarray, L_full: np.ndarray, A_inv_u: np.ndarray, B: np.ndarray, v: np.ndarray) -> np.ndarray:
    """``BlockMatrix2x2.schur_update`` (_block.py:226-231): y = S^-1 (v - C A^-1 u); x = A^-1 u - A^-1 B y."""
    n0 = L_A.shape[0]
    L_S = L_full[n0:, n0:]
    y = cho_solve_lower(L_S, v - B.T @ A_inv_u)
    x = A_inv_u - cho_solve_lower(L_A, B @ y)
    return np.concatenate((x, y))
