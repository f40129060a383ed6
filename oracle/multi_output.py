"""Oracle: conditioning of a MULTI-OUTPUT GP with independent outputs (numpy restatement).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows
src/linpde_gp/randprocs/covfuncs/_independent_multi_output.py:11-49 (block-diagonal prior covariance),
src/linpde_gp/randprocs/covfuncs/_stack.py:15-104, src/linpde_gp/linfuncops/_select_output.py:9-34 and the
``SelectOutput`` registrations src/linpde_gp/randprocs/covfuncs/linfuncops/_registry.py:34-48, 82-120: an
observation operator is a sum ``L = sum_t c_t * (D_t @ SelectOutput(o_t))`` with scalar-output differential
operators ``D_t`` (or the identity), so that

    (L_a k L_b^*)(x, x') = sum_{s, t : o_s == o_t} c_s c_t (D_s k_{o_s} D_t^*)(x, x')

and ``SelectOutput(j)`` of the posterior has cross-covariance ``sum_{t : o_t == j} c_t (k_j D_t^*)(x, X_b)`` with
observation block ``b`` (experiments/0000_cpu_stationary_1d.ipynb cells 55-82: joint belief over temperature, volumetric
and surface heat sources).

A *problem* is a JSON-able dict:
    {"kernels": [<kernel spec> per output], "means": [float per output],
     "blocks": [{"X": ..., "Y": ..., "Ls": [[output, scalar, op | None], ...], "noise_var": ...}, ...],
     "Xt": (M, d) list, "n_cov": int}
"""
from __future__ import annotations

import numpy as np

from . import covfuncs as ocf
from . import gp as ogp
from . import linalg as ola


def _terms(Ls):
    return [(int(o), float(c), ogp._op(op) if op is not None else None) for o, c, op in Ls]  # pylint: disable=protected-access


def block(kernels, La, Lb, Xa, Xb=None):
    """(L_a k L_b^*)(X_a, X_b) for the block-diagonal multi-output kernel; ``Xb=None`` -> ``Xb := Xa``."""
    Xa = np.asarray(Xa, dtype=np.double)
    out = None
    for oa, ca, opa in La:
        for ob, cb, opb in Lb:
            if oa != ob:
                continue
            v = (ca * cb) * ocf.matrix(kernels[oa], opa, opb, Xa, Xb)
            out = v if out is None else out + v
    if out is None:
        n0 = ocf.matrix(kernels[0], None, None, Xa, Xb).shape
        out = np.zeros(n0)
    return out


class Posterior:
    def __init__(self, kernels, means):
        self.kernels, self.means = kernels, [float(m) for m in means]
        self.Xs, self.Ls = [], []
        self.L = np.zeros((0, 0))
        self.resid = np.zeros((0,))
        self.gram = np.zeros((0, 0))
        self.w = np.zeros((0,))

    def condition(self, Y, X, Ls, noise_var=None):
        X = np.asarray(X, dtype=np.double)
        Y = np.asarray(Y, dtype=np.double).reshape(-1)
        T = _terms(Ls)
        # L applied to the constant prior mean: only order-zero parts survive (functions/_constant.py)
        pred = np.zeros_like(Y)
        for o, c, op in T:
            if op is None:
                pred = pred + c * self.means[o]
        D = block(self.kernels, T, T, X, None)
        if noise_var is not None:
            D = D + np.diag(np.broadcast_to(np.asarray(noise_var, dtype=np.double), Y.shape))
        new = Posterior(self.kernels, self.means)
        new.Xs, new.Ls = self.Xs + [X], self.Ls + [T]
        new.resid = np.concatenate([self.resid, Y - pred])
        if self.L.shape[0] == 0:
            new.L = ola.cholesky_lower(D)
            new.gram = D
        else:
            C = np.concatenate([block(self.kernels, T, Lj, X, Xj) for Xj, Lj in zip(self.Xs, self.Ls)], axis=1)
            new.L = ola.cholesky_append(self.L, C.T, D)
            new.gram = np.block([[self.gram, C.T], [C, D]])
        new.w = ola.cho_solve_lower(new.L, new.resid)
        return new

    def crosscov(self, j, Xt):
        sel = [(j, 1.0, None)]
        return np.concatenate([block(self.kernels, sel, Lb, Xt, Xb) for Xb, Lb in zip(self.Xs, self.Ls)], axis=1)

    def mean(self, j, Xt):
        return self.means[j] + self.crosscov(j, Xt) @ self.w

    def var(self, j, Xt):
        V = ola.solve_lower(self.L, self.crosscov(j, Xt).T)
        return ocf.diagonal(self.kernels[j], None, None, Xt) - np.sum(V * V, axis=0)

    def cov(self, j, Xt):
        K = self.crosscov(j, Xt)
        return ocf.matrix(self.kernels[j], None, None, Xt, None) - K @ ola.cho_solve_lower(self.L, K.T)


def solve(problem):
    """Posterior of every selected output: ``mean`` / ``var`` have shape (n_outputs, M), ``cov`` (n_outputs, c, c)."""
    post = Posterior(problem["kernels"], problem["means"])
    for blk in problem["blocks"]:
        post = post.condition(blk["Y"], blk["X"], blk["Ls"], blk.get("noise_var"))
    Xt = np.asarray(problem["Xt"], dtype=np.double)
    Xc = Xt[: problem.get("n_cov", 8)]
    nout = len(problem["kernels"])
    return {
        "w": post.w,
        "gram": post.gram,
        "mean": np.stack([post.mean(j, Xt) for j in range(nout)]),
        "var": np.stack([post.var(j, Xt) for j in range(nout)]),
        "cov": np.stack([post.cov(j, Xc) for j in range(nout)]),
    }


# ----------------------------------------------------------------------------------------
# golden problems
# ----------------------------------------------------------------------------------------
def _k(scale, base):
    return {"scale": float(scale), "base": base}


def cpu_1d_problem(n_pde=17, n_dts=5, grid=25):
    """The joint (u, q_V, q_A) model of experiments/0000_cpu_stationary_1d.ipynb cells 55-82 without the stationarity
    functional: PDE ``-kappa u'' - q_V = 0``, Neumann ``-kappa d_n u - q_A = 0`` at both ends, noisy temperature
    readings ``u(x_i) = y_i``."""
    m = ogp._m  # pylint: disable=protected-access
    kappa = 0.7
    rng = np.random.default_rng(11)
    X = np.linspace(0.03, 0.97, n_pde)
    Xd = np.sort(rng.uniform(0.1, 0.9, n_dts))
    yd = 58.0 + rng.normal(size=n_dts)
    lap = [[-kappa, ["wl", 1.0]]]
    blocks = [
        {"X": X.tolist(), "Y": np.zeros(n_pde).tolist(), "Ls": [[0, 1.0, lap], [1, -1.0, None]], "noise_var": None},
        {"X": [0.0], "Y": [0.0], "Ls": [[0, 1.0, [[-kappa, ["dd", -1.0]]]], [2, -1.0, None]], "noise_var": None},
        {"X": [1.0], "Y": [0.0], "Ls": [[0, 1.0, [[-kappa, ["dd", 1.0]]]], [2, -1.0, None]], "noise_var": None},
        {"X": Xd.tolist(), "Y": yd.tolist(), "Ls": [[0, 1.0, None]], "noise_var": 0.25},
    ]
    return {
        "kernels": [_k(9.0, m(2.5, 0.75)), _k(0.81, m(0.5, 1.0)), _k(0.81, m(0.5, 1.0))],
        "means": [57.0, 0.3, -0.1],
        "blocks": blocks,
        "Xt": np.linspace(0.0, 1.0, grid).tolist(),
        "n_cov": 8,
    }


def poisson2d_joint_problem(n_pde=60, n_bc_edge=6, n_f=12, grid=7, seed=3):
    """2-D Poisson problem with an UNCERTAIN right-hand side modelled jointly: outputs (u, f), ``-Δu - f = 0`` at
    collocation points, Dirichlet data for u, noisy point measurements of f."""
    tp, m = ogp._tp, ogp._m  # pylint: disable=protected-access
    rng = np.random.default_rng(seed)
    Xp = rng.uniform(0.0, 1.0, size=(n_pde, 2))
    t = np.linspace(0.0, 1.0, n_bc_edge, endpoint=False)
    Xb = np.concatenate([np.stack([t, 0 * t], 1), np.stack([1 + 0 * t, t], 1), np.stack([1 - t, 1 + 0 * t], 1),
                         np.stack([0 * t, 1 - t], 1)])
    Xf = rng.uniform(0.0, 1.0, size=(n_f, 2))
    yf = 2.0 + 0.3 * rng.normal(size=n_f)
    g = np.linspace(0.0, 1.0, grid)
    Xt = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    blocks = [
        {"X": Xp.tolist(), "Y": np.zeros(n_pde).tolist(), "Ls": [[0, 1.0, [[-1.0, ["wl", [1.0, 1.0]]]]], [1, -1.0, None]],
         "noise_var": None},
        {"X": Xb.tolist(), "Y": np.zeros(len(Xb)).tolist(), "Ls": [[0, 1.0, None]], "noise_var": None},
        {"X": Xf.tolist(), "Y": yf.tolist(), "Ls": [[1, 1.0, None]], "noise_var": 0.01},
    ]
    return {
        "kernels": [_k(4.0, tp(m(2.5, 0.4), m(2.5, 0.4))), _k(1.0, tp(m(1.5, 0.5), m(1.5, 0.5)))],
        "means": [0.0, 2.0],
        "blocks": blocks,
        "Xt": Xt.tolist(),
        "n_cov": 8,
    }


def golden_problems():
    return {"cpu_1d": cpu_1d_problem(), "poisson2d_joint": poisson2d_joint_problem()}
