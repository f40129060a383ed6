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

Scalar FUNCTIONAL observations (second half): a block may instead carry
``"functional": [["int", output, coef, a, b] | ["eval", output, coef, x, op | None], ...]`` -- one row that is the sum
of Lebesgue integrals ``coef * int_a^b f_output`` (closed forms of ``oracle/integrals.py``) and point evaluations, the
stationarity condition ``h * LebesgueIntegral(domain) @ select_q_V + h * (select_q_A.to_linfunctl(w) +
select_q_A.to_linfunctl(0))`` of experiments/0000_cpu_stationary_1d.ipynb cells 65-66, 85.

A *problem* is a JSON-able dict:
    {"kernels": [<kernel spec> per output], "means": [float per output],
     "blocks": [{"X": ..., "Y": ..., "Ls": [[output, scalar, op | None], ...], "noise_var": ...}
                | {"functional": [...], "Y": [y]}, ...],
     "Xt": (M, d) list, "n_cov": int}
"""
from __future__ import annotations

import numpy as np

from . import covfuncs as ocf
from . import gp as ogp
from . import integrals as oint
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


# An observation batch is a SUM OF ATOMS, every atom producing the same number of rows:
#   ("pts", output, coef, op | None, X)      coef * (op f_output)(X)          -- point evaluations
#   ("int", output, coef, (a, b))            coef * int_a^b f_output(t) dt    -- one row (LebesgueIntegral @ SelectOutput)
# (src/linpde_gp/linfunctls/_arithmetic.py:12-90 Scaled / Sum functionals, :92-140 CompositeLinearFunctional).
def _atoms_of(blk):
    if "functional" not in blk:
        X = np.asarray(blk["X"], dtype=np.double)
        return [("pts", o, c, op, X) for o, c, op in _terms(blk["Ls"])]
    atoms = []
    for a in blk["functional"]:
        if a[0] == "int":
            atoms.append(("int", int(a[1]), float(a[2]), (float(a[3]), float(a[4]))))
        else:  # ["eval", output, coef, x, op | None]
            op = ogp._op(a[4]) if len(a) > 4 and a[4] is not None else None  # pylint: disable=protected-access
            atoms.append(("pts", int(a[1]), float(a[2]), op, np.asarray([a[3]], dtype=np.double)))
    return atoms


def _rows(atoms):
    return 1 if atoms[0][0] == "int" else len(atoms[0][4])


def _atom_cov(kernels, A, B):
    """cov(atom A, atom B) for independent outputs (zero across outputs)."""
    k = kernels[A[1]]
    cc = A[2] * B[2]
    if A[0] == "pts" and B[0] == "pts":
        return cc * ocf.matrix(k, A[3], B[3], A[4], B[4])
    if A[0] == "int" and B[0] == "int":
        return cc * np.full((1, 1), oint.integral_integral(k, A[3], B[3]))
    pts, itg = (A, B) if A[0] == "pts" else (B, A)
    if pts[3] is not None:
        raise NotImplementedError("a differential operator applied to an integral cross-covariance (not in the reference either)")
    v = cc * oint.integral_crosscov(k, itg[3], pts[4]).reshape(-1, 1)
    return v if A[0] == "pts" else v.T


def atoms_cov(kernels, atoms_a, atoms_b):
    out = np.zeros((_rows(atoms_a), _rows(atoms_b)))
    for A in atoms_a:
        for B in atoms_b:
            if A[1] == B[1]:
                out = out + _atom_cov(kernels, A, B)
    return out


def _atom_prior_mean(means, A):
    """L applied to the constant prior mean: order-zero parts (functions/_constant.py), integrals -> value * volume
    (src/linpde_gp/linfunctls/_integrals.py:60-62)."""
    if A[0] == "int":
        return A[2] * means[A[1]] * (A[3][1] - A[3][0])
    return A[2] * means[A[1]] if A[3] is None else 0.0


class Posterior:
    def __init__(self, kernels, means):
        self.kernels, self.means = kernels, [float(m) for m in means]
        self.blocks = []
        self.L = np.zeros((0, 0))
        self.resid = np.zeros((0,))
        self.gram = np.zeros((0, 0))
        self.w = np.zeros((0,))

    def condition(self, blk):
        atoms = _atoms_of(blk)
        Y = np.asarray(blk["Y"], dtype=np.double).reshape(-1)
        noise_var = blk.get("noise_var")
        pred = np.zeros_like(Y)
        for A in atoms:
            pred = pred + _atom_prior_mean(self.means, A)
        D = atoms_cov(self.kernels, atoms, atoms)
        if noise_var is not None:
            D = D + np.diag(np.broadcast_to(np.asarray(noise_var, dtype=np.double), Y.shape))
        new = Posterior(self.kernels, self.means)
        new.blocks = self.blocks + [atoms]
        new.resid = np.concatenate([self.resid, Y - pred])
        if self.L.shape[0] == 0:
            new.L = ola.cholesky_lower(D)
            new.gram = D
        else:
            C = np.concatenate([atoms_cov(self.kernels, atoms, prev) for prev in self.blocks], axis=1)
            new.L = ola.cholesky_append(self.L, C.T, D)
            new.gram = np.block([[self.gram, C.T], [C, D]])
        new.w = ola.cho_solve_lower(new.L, new.resid)
        return new

    def crosscov(self, j, Xt):
        sel = [("pts", j, 1.0, None, np.asarray(Xt, dtype=np.double))]
        return np.concatenate([atoms_cov(self.kernels, sel, b) for b in self.blocks], axis=1)

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
        post = post.condition(blk)
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


def _stationarity_block(h=0.4, width=1.0):
    """``h * int_0^w q_V + h * (q_A(w) + q_A(0)) = 0`` (notebook cell 65)."""
    return {"functional": [["int", 1, h, 0.0, width], ["eval", 2, h, width, None], ["eval", 2, h, 0.0, None]], "Y": [0.0]}


def cpu_1d_stat_problem():
    """cpu_1d plus the stationarity functional as the LAST observation (notebook cell 85); q_A is Matern-3/2 here so
    that two different antiderivative polynomials are exercised."""
    prob = cpu_1d_problem()
    prob["kernels"][1] = _k(0.81, ogp._m(1.5, 0.8))  # pylint: disable=protected-access
    prob["blocks"] = prob["blocks"] + [_stationarity_block()]
    return prob


def cpu_1d_stat_first_problem():
    """Stationarity functional FIRST (notebook cell 66), then PDE and temperature observations appended to it."""
    prob = cpu_1d_problem(n_pde=9, n_dts=4, grid=17)
    prob["blocks"] = [_stationarity_block(h=0.25)] + prob["blocks"]
    return prob


def golden_problems():
    return {"cpu_1d": cpu_1d_problem(), "poisson2d_joint": poisson2d_joint_problem(),
            "cpu_1d_stat": cpu_1d_stat_problem(), "cpu_1d_stat_first": cpu_1d_stat_first_problem()}
