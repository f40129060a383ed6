"""Oracle: GP conditioning on linear observations and posterior evaluation (numpy restatement).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows
src/linpde_gp/randprocs/_gaussian_process/_conditional.py: ``from_observations`` :27-54,
``condition_on_observations`` :253-294 (block append through ``BlockMatrix2x2.schur_update``),
``Mean._evaluate`` :193-197, ``CovarianceFunction._evaluate`` :223-231 / ``_evaluate_linop`` :245-251,
``_preprocess_observations`` :296-399 (noise added to the Gram, :392-394).  Zero prior mean.

A *problem* is a JSON-able dict:
    {"kernel": <kernel spec>, "blocks": [{"X": (N_i, d) list, "Y": (N_i,) list, "L": op | None,
     "noise_var": None | float | (N_i,) list}, ...], "Xt": (M, d) list, "n_cov": int}
"""
from __future__ import annotations

import numpy as np

from . import covfuncs as ocf
from . import linalg as ola


def _op(L):
    if L is None:
        return None
    out = []
    for scalar, (kind, payload) in L:
        if kind == "pd":
            out.append((scalar, ("pd", {tuple(mi): c for mi, c in payload})))
        else:
            out.append((scalar, (kind, np.asarray(payload, dtype=float))))
    return out


class Posterior:
    """State of a ``ConditionalGaussianProcess`` (_conditional.py:67-77): observation blocks, the lower
    Cholesky factor of the Gram matrix (grown by bordering) and the representer weights."""

    def __init__(self, kernel):
        self.kernel = kernel
        self.Xs, self.Ls, self.Ys = [], [], []
        self.L = np.zeros((0, 0))
        self.w = np.zeros((0,))
        self.gram = np.zeros((0, 0))

    def condition(self, Y, X, L=None, noise_var=None):
        X = np.asarray(X, dtype=np.double)
        Y = np.asarray(Y, dtype=np.double).reshape(-1)
        op = _op(L) if (L is not None and not isinstance(L[0], tuple)) else L
        D = ocf.matrix(self.kernel, op, op, X, None)
        if noise_var is not None:
            D = D + np.diag(np.broadcast_to(np.asarray(noise_var, dtype=np.double), Y.shape))
        n0 = self.L.shape[0]
        new = Posterior(self.kernel)
        new.Xs, new.Ls, new.Ys = self.Xs + [X], self.Ls + [op], self.Ys + [Y]
        if n0 == 0:
            new.L = ola.cholesky_lower(D)
            new.w = ola.cho_solve_lower(new.L, Y)
            new.gram = D
            return new
        # lower-left blocks L_new k L_j^*(X_new, X_j)  (_conditional.py:270, :420-429)
        C = np.concatenate([ocf.matrix(self.kernel, op, Lj, X, Xj) for Xj, Lj in zip(self.Xs, self.Ls)], axis=1)
        B = C.T
        new.L = ola.cholesky_append(self.L, B, D)
        new.w = ola.schur_update(self.L, new.L, self.w, B, Y)
        new.gram = np.block([[self.gram, B], [C, D]])
        return new

    def crosscov(self, Xt, Lt=None):
        """(Lt k L_j^*)(Xt, X_j) concatenated over blocks (``PriorPredictiveCrossCovariance._evaluate`` :140-153)."""
        Xt = np.asarray(Xt, dtype=np.double)
        return np.concatenate([ocf.matrix(self.kernel, Lt, Lj, Xt, Xj) for Xj, Lj in zip(self.Xs, self.Ls)], axis=1)

    def mean(self, Xt):
        return self.crosscov(Xt) @ self.w

    def var(self, Xt):
        K = self.crosscov(Xt)
        V = ola.solve_lower(self.L, K.T)
        return ocf.diagonal(self.kernel, None, None, Xt) - np.sum(V * V, axis=0)

    def cov(self, Xt0, Xt1=None):
        K0 = self.crosscov(Xt0)
        K1 = K0 if Xt1 is None else self.crosscov(Xt1)
        return ocf.matrix(self.kernel, None, None, Xt0, Xt1) - K0 @ ola.cho_solve_lower(self.L, K1.T)


def solve(problem):
    post = Posterior(problem["kernel"])
    for blk in problem["blocks"]:
        post = post.condition(blk["Y"], blk["X"], blk["L"], blk.get("noise_var"))
    Xt = np.asarray(problem["Xt"], dtype=np.double)
    return {
        "w": post.w,
        "mean": post.mean(Xt),
        "var": post.var(Xt),
        "cov": post.cov(Xt[: problem.get("n_cov", 16)]),
        "gram": post.gram,
    }


# ----------------------------------------------------------------------------------------
# small golden problems (shapes follow the reference's experiments / tests)
# ----------------------------------------------------------------------------------------
def _tp(*f):
    return {"kind": "tensor_product", "factors": list(f)}


def _m(nu, ell):
    return {"kind": "matern", "input_shape": [], "nu": float(nu), "lengthscales": float(ell)}


def _neg_lap(d):
    return [[-1.0, ["wl", [1.0] * d if d else 1.0]]]


def _heat(alpha):
    return [[1.0, ["pd", [[[1, 0], 1.0]]]], [1.0, ["wl", [0.0, -float(alpha)]]]]


def poisson2d_problem(n_pde, n_bc_edge, ell=None, seed=0, grid=12, sigma2=4.0, nu=2.5, noise_bc=None, lo=0.0, hi=1.0):
    """2-D Poisson Dirichlet problem -Δu = 2, u|∂Ω = 0 on [lo,hi]^2 with a product-Matern prior
    (experiments/0001_poisson_dirichlet_2d.ipynb cells 9-19): 4 boundary blocks then
    the PDE block."""
    rng = np.random.default_rng(seed)
    ell = 4.0 / np.sqrt(n_pde) * (hi - lo) if ell is None else ell
    kernel = {"scale": sigma2, "base": _tp(_m(nu, ell), _m(nu, ell))}
    s = np.linspace(lo, hi, n_bc_edge, endpoint=False)
    edges = [
        np.stack([s, np.full_like(s, lo)], -1),
        np.stack([np.full_like(s, hi), s], -1),
        np.stack([s + (hi - lo) / n_bc_edge, np.full_like(s, hi)], -1),
        np.stack([np.full_like(s, lo), s + (hi - lo) / n_bc_edge], -1),
    ]
    blocks = [{"X": e.tolist(), "Y": np.zeros(len(e)).tolist(), "L": None, "noise_var": noise_bc} for e in edges]
    Xp = rng.uniform(lo, hi, size=(n_pde, 2))
    blocks.append({"X": Xp.tolist(), "Y": np.full(n_pde, 2.0).tolist(), "L": _neg_lap(2), "noise_var": None})
    g = np.linspace(lo, hi, grid)
    Xt = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    return {"kernel": kernel, "blocks": blocks, "Xt": Xt.tolist(), "n_cov": 16}


def heat_problem(n_ic=5, n_bc=10, nt=15, nx=8, alpha=0.1, grid=10):
    """1-D heat equation IBVP (tests/linpde_gp/problems/test_heat.py:56-99): prior
    TensorProduct(Matern-3/2(t, 2.5), Matern-5/2(x, 2.0)), IC / two noisy BC blocks / PDE block."""
    kernel = {"scale": None, "base": _tp(_m(1.5, 2.5), _m(2.5, 2.0))}
    xs = np.linspace(-1.0, 1.0, n_ic)
    X_ic = np.stack([np.zeros_like(xs), xs], -1)
    Y_ic = np.sin(np.pi * (xs + 1.0) / 2.0)
    ts = np.linspace(0.0, 5.0, n_bc)
    blocks = [{"X": X_ic.tolist(), "Y": Y_ic.tolist(), "L": None, "noise_var": None}]
    for xb in (-1.0, 1.0):
        Xb = np.stack([ts, np.full_like(ts, xb)], -1)
        blocks.append({"X": Xb.tolist(), "Y": np.zeros(n_bc).tolist(), "L": None, "noise_var": 1e-5})
    tg = np.linspace(0.0, 5.0, nt)
    xg = np.linspace(-1.0, 1.0, nx + 2)[1:-1]
    Xp = np.stack(np.meshgrid(tg, xg, indexing="ij"), -1).reshape(-1, 2)
    blocks.append({"X": Xp.tolist(), "Y": np.zeros(len(Xp)).tolist(), "L": _heat(alpha), "noise_var": None})
    Xt = np.stack(np.meshgrid(np.linspace(0, 5, grid), np.linspace(-1, 1, grid), indexing="ij"), -1).reshape(-1, 2)
    return {"kernel": kernel, "blocks": blocks, "Xt": Xt.tolist(), "n_cov": 16}


def _grid_block(factors, Y, L, noise_var=None):
    """Observation batch on a tensor-product grid: ``X`` is the C-order flattening (what the reference sees after
    ``np.asarray``), ``grid`` keeps the 1-D factors for implementations that exploit the Kronecker structure."""
    X = np.stack(np.meshgrid(*[np.asarray(f, dtype=float) for f in factors], indexing="ij"), -1).reshape(-1, len(factors))
    Y = np.broadcast_to(np.asarray(Y, dtype=float), (len(X),))
    return {"X": X.tolist(), "Y": Y.tolist(), "L": L, "noise_var": noise_var, "grid": [np.asarray(f, dtype=float).tolist() for f in factors]}


def poisson2d_grid_problem(nx=9, ny=8, n_bc_edge=10, ell=0.35, sigma2=4.0, grid=10):
    """As :func:`poisson2d_problem`, with ALL batches on tensor-product grids (boundary edges = grids with one
    singleton factor): every Gram block is a sum of Kronecker products."""
    kernel = {"scale": sigma2, "base": _tp(_m(2.5, ell), _m(2.5, ell))}
    s = np.linspace(0.0, 1.0, n_bc_edge)
    blocks = [
        _grid_block([s, [0.0]], 0.0, None),
        _grid_block([[1.0], s[1:]], 0.0, None),
        _grid_block([s[:-1], [1.0]], 0.0, None),
        _grid_block([[0.0], s[1:-1]], 0.0, None),
        _grid_block([np.linspace(0.0, 1.0, nx + 2)[1:-1], np.linspace(0.0, 1.0, ny + 2)[1:-1]], 2.0, _neg_lap(2)),
    ]
    g = np.linspace(0.0, 1.0, grid)
    Xt = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    return {"kernel": kernel, "blocks": blocks, "Xt": Xt.tolist(), "n_cov": 16}


def heat_grid_problem(n_ic=7, n_bc=9, nt=11, nx=6, alpha=0.1, grid=10):
    """As :func:`heat_problem` with every batch given as a tensor-product grid."""
    kernel = {"scale": None, "base": _tp(_m(1.5, 2.5), _m(2.5, 2.0))}
    xs = np.linspace(-1.0, 1.0, n_ic)
    ts = np.linspace(0.0, 5.0, n_bc)[1:]
    ic = _grid_block([[0.0], xs], 0.0, None)
    ic["Y"] = np.sin(np.pi * (xs + 1.0) / 2.0).tolist()
    blocks = [ic, _grid_block([ts, [-1.0]], 0.0, None, 1e-5), _grid_block([ts, [1.0]], 0.0, None, 1e-5),
              _grid_block([np.linspace(0.0, 5.0, nt), np.linspace(-1.0, 1.0, nx + 2)[1:-1]], 0.0, _heat(alpha))]
    Xt = np.stack(np.meshgrid(np.linspace(0, 5, grid), np.linspace(-1, 1, grid), indexing="ij"), -1).reshape(-1, 2)
    return {"kernel": kernel, "blocks": blocks, "Xt": Xt.tolist(), "n_cov": 16}


def golden_problems():
    probs = {}
    probs["poisson2d_grid_kron"] = poisson2d_grid_problem()
    probs["heat_grid_kron"] = heat_grid_problem()
    # C1: experiments/0000_poisson_dirichlet_1d.ipynb (cells 9, 17, 21): PDE first, then boundary
    xp = np.linspace(-0.8, 0.8, 3)
    probs["poisson1d_expquad"] = {
        "kernel": {"scale": 4.0, "base": {"kind": "expquad", "input_shape": [], "lengthscales": 1.0}},
        "blocks": [
            {"X": xp.tolist(), "Y": [2.0] * 3, "L": _neg_lap(0), "noise_var": None},
            {"X": [-1.0, 1.0], "Y": [0.0, 0.0], "L": None, "noise_var": None},
        ],
        "Xt": np.linspace(-1, 1, 100).tolist(),
        "n_cov": 16,
    }
    xp = np.linspace(-0.8, 0.8, 40)
    probs["poisson1d_matern35"] = {
        "kernel": {"scale": 4.0, "base": {"kind": "matern", "input_shape": [], "nu": 3.5, "lengthscales": 0.5}},
        "blocks": [
            {"X": [-1.0, 1.0], "Y": [0.0, 0.0], "L": None, "noise_var": None},
            {"X": xp.tolist(), "Y": (np.pi**2 * np.sin(np.pi * xp)).tolist(), "L": _neg_lap(0), "noise_var": None},
        ],
        "Xt": np.linspace(-1, 1, 100).tolist(),
        "n_cov": 16,
    }
    probs["poisson2d_tp_matern25"] = poisson2d_problem(200, 16, seed=0, grid=12)
    probs["poisson2d_tp_matern25_noisybc"] = poisson2d_problem(150, 12, seed=3, grid=10, noise_bc=1e-6, lo=-1.0, hi=1.0)
    probs["heat_tp_matern"] = heat_problem()
    # tests/linpde_gp/randprocs/test_posterior_gp.py:25-88: ExpQuad, 11 points, batches (2,3,2,4), two noisy
    rng = np.random.default_rng(25)
    xs = np.linspace(-1.0, 1.0, 11)
    ys = np.sin(3 * xs) + 0.1 * rng.standard_normal(11)
    blocks, start = [], 0
    for size, nv in zip((2, 3, 2, 4), (None, 0.6**2, None, 0.3**2)):
        blocks.append({"X": xs[start : start + size].tolist(), "Y": ys[start : start + size].tolist(), "L": None, "noise_var": nv})
        start += size
    probs["expquad_iterative"] = {
        "kernel": {"scale": 4.0, "base": {"kind": "expquad", "input_shape": [], "lengthscales": 0.25}},
        "blocks": blocks,
        "Xt": np.linspace(-1.5, 1.5, 50).tolist(),
        "n_cov": 16,
    }
    rng = np.random.default_rng(7)
    Xp = rng.uniform(-1, 1, size=(60, 2))
    th = np.linspace(0, 2 * np.pi, 24, endpoint=False)
    Xb = np.stack([np.cos(th), np.sin(th)], -1)
    probs["poisson2d_expquad_ard"] = {
        "kernel": {"scale": 2.25, "base": {"kind": "expquad", "input_shape": [2], "lengthscales": [0.25, 0.35]}},
        "blocks": [
            {"X": Xb.tolist(), "Y": np.zeros(24).tolist(), "L": None, "noise_var": 1e-6},
            {"X": Xp.tolist(), "Y": np.full(60, 2.0).tolist(), "L": _neg_lap(2), "noise_var": 1e-6},
        ],
        "Xt": rng.uniform(-0.7, 0.7, size=(40, 2)).tolist(),
        "n_cov": 16,
    }
    return probs
