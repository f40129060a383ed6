"""Parity at BASELINE.json's full sizes through size-independent properties (the CPU oracle cannot run there).

For a GP conditioned on noise-free linear observations  L_i u(X_i) = y_i  the posterior has exact, checkable
properties that hold at any N (src/linpde_gp/randprocs/_gaussian_process/_conditional.py:193-251):
  * interpolation: the posterior mean reproduces every observation, L_i m(X_i) = y_i, and the posterior variance of
    an observed quantity vanishes;
  * 0 <= var(x) <= k(x, x); the posterior covariance matrix is symmetric positive semi-definite;
  * the cached factor solves the Gram system: with G v evaluated MATRIX-FREE by the posterior-mean kernel (an
    independent code path that never sees the assembled matrix), potrs(G v) = v;
  * linearity of the representer weights in the observations.
All of it runs through the public API / the C ABI on the configs C2 (N = 16,384), C3 (heat, N = 32,768) and C4
(N = 65,536); tolerances are the north-star ones (1e-8 on posterior quantities, relative to the problem's scale).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

POST_TOL = 1e-8


def _poisson(n_pde, n_bc_edge):
    import bench
    import linpde_gp_b200 as lg
    from linpde_gp_b200.linfuncops import diffops
    from linpde_gp_b200.randprocs import covfuncs

    prob = bench.make_problem(n_pde, n_bc_edge, 64)
    k = bench.SIGMA2 * covfuncs.TensorProduct(covfuncs.Matern((), nu=bench.NU, lengthscales=prob["ell"]),
                                              covfuncs.Matern((), nu=bench.NU, lengthscales=prob["ell"]))
    prior = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), k)
    L = -1.0 * diffops.Laplacian((2,))
    batches = [(Yb, Xb, None) for Xb, Yb in zip(prob["edges"], prob["Y_bc"])] + [(prob["Y_pde"], prob["X_pde"], L)]
    return prior, batches, prob["Xt"], bench.SIGMA2


def _heat(n_ic, n_bc, n_pde):
    import linpde_gp_b200 as lg
    from linpde_gp_b200.linfuncops import diffops
    from linpde_gp_b200.randprocs import covfuncs

    rng = np.random.default_rng(1)
    # lengthscales scaled with the point density per axis (equilibrated cond ~1e5, as for C2/C4)
    lt, lx = 5.0 * 2.0 / np.sqrt(n_pde), 2.0 * 2.0 / np.sqrt(n_pde)
    k = covfuncs.TensorProduct(covfuncs.Matern((), nu=1.5, lengthscales=lt), covfuncs.Matern((), nu=2.5, lengthscales=lx))
    prior = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), k)
    xs = np.linspace(-1.0, 1.0, n_ic)
    X_ic = np.stack([np.zeros_like(xs), xs], -1)
    ts = np.linspace(0.0, 5.0, n_bc + 1)[1:]
    batches = [(np.sin(np.pi * (xs + 1.0) / 2.0), X_ic, None)]
    for xb in (-1.0, 1.0):
        batches.append((np.zeros(n_bc), np.stack([ts, np.full_like(ts, xb)], -1), None))
    Xp = np.stack([rng.uniform(0.0, 5.0, n_pde), rng.uniform(-1.0, 1.0, n_pde)], -1)
    batches.append((np.zeros(n_pde), Xp, diffops.HeatOperator(domain_shape=(2,), alpha=0.1)))
    Xt = np.stack(np.meshgrid(np.linspace(0, 5, 64), np.linspace(-1, 1, 64), indexing="ij"), -1).reshape(-1, 2)
    return prior, batches, Xt, 1.0


def _check_posterior_properties(prior, batches, Xt, sigma2, one_shot):
    import linpde_gp_b200 as lg
    from linpde_gp_b200 import backend

    if one_shot:
        post = lg.ConditionalGaussianProcess.from_observation_batches(prior, batches)
    else:
        post = prior
        for Y, X, L in batches:
            post = post.condition_on_observations(Y, X=X, L=L)
    rng = np.random.default_rng(0)
    # (1) interpolation of every batch (a random subset of 512 points each), variance of observed quantities = 0
    for Y, X, L in batches:
        idx = rng.choice(len(X), size=min(512, len(X)), replace=False)
        gp = post if L is None else L(post)
        prior_scale = float(np.max(np.abs((prior if L is None else L(prior)).var(X[idx[:8]]))))
        m = gp.mean(X[idx])
        assert np.max(np.abs(m - Y[idx])) <= POST_TOL * max(1.0, np.max(np.abs(Y)), np.sqrt(prior_scale)), (L, np.max(np.abs(m - Y[idx])))
        v = gp.var(X[idx[:128]])
        assert np.max(np.abs(v)) <= POST_TOL * prior_scale, (L, np.max(np.abs(v)), prior_scale)
    # (2) variance bounds and a PSD covariance block on the test grid
    var = post.var(Xt)
    assert var.shape == (len(Xt),) and np.all(np.isfinite(var))
    assert var.min() >= -POST_TOL * sigma2 and var.max() <= sigma2 * (1.0 + POST_TOL)
    C = post.cov.matrix(Xt[:256])
    assert np.max(np.abs(C - C.T)) <= POST_TOL * sigma2
    assert np.linalg.eigvalsh(0.5 * (C + C.T)).min() >= -POST_TOL * sigma2
    assert np.max(np.abs(np.diag(C) - var[:256])) <= POST_TOL * sigma2
    # (3) factor vs matrix-free Gram products:  potrs(G v) == v, G v from the posterior-mean kernel, row block by
    #     row block with the operator of that block on the test side
    fac = post._factor
    n = fac.n
    #     (v and the error are measured in the equilibrated variables D v, D = sqrt(diag G): the Gram matrix mixes
    #     k entries of size sigma^2 with L k L* entries ~1e9 times larger, SURVEY 8d quotes the equilibrated condition)
    dscale = torch.ones(n, dtype=torch.float64, device=fac.L.device)
    for blk in post._blocks:
        kj = prior.cov if blk.op is None else blk.op(prior.cov, argnum=1)
        kii = kj if blk.op is None else blk.op(kj, argnum=0)
        dscale[blk.col_off : blk.col_off + blk.n] = float(np.sqrt(kii.descriptor().diag_value))
    z = torch.from_numpy(rng.standard_normal(n)).to(fac.L.device)
    v = z / dscale
    Gv = torch.empty(n, dtype=torch.float64, device=v.device)
    for blk in post._blocks:
        cols_d, cols_X, cols_off = [], [], []
        for pb in post._blocks:
            kj = prior.cov if pb.op is None else pb.op(prior.cov, argnum=1)
            kij = kj if blk.op is None else blk.op(kj, argnum=0)
            cols_d.append(kij.descriptor())
            cols_X.append(pb.X)
            cols_off.append(pb.col_off)
        backend.post_mean(backend.ObsBlocks(cols_d, cols_X, cols_off), v, blk.X, out=Gv[blk.col_off : blk.col_off + blk.n])
    x = fac.potrs(Gv.clone().reshape(1, -1)).reshape(-1)
    rel = float(((x - v) * dscale).abs().max() / z.abs().max())
    assert rel <= POST_TOL, rel
    # (4) linearity of the representer weights:  w(y) for the stacked observations y equals potrs(y)
    y = torch.cat([backend.to_device(np.asarray(Y, dtype=float)) for Y, _, _ in batches])
    w = fac.potrs(y.clone().reshape(1, -1)).reshape(-1)
    w2 = fac.potrs((3.0 * y + Gv).reshape(1, -1).clone()).reshape(-1)
    assert float(((w2 - (3.0 * w + x)) * dscale).abs().max()) <= POST_TOL * float((w2 * dscale).abs().max())
    assert float((w - post._w).abs().max()) <= 1e-12 * float(w.abs().max())
    assert np.isfinite(fac.logdet())


def test_c2_poisson_n16384_properties():
    _check_posterior_properties(*_poisson(15360, 256), one_shot=False)


def test_c3_heat_n32768_properties():
    _check_posterior_properties(*_heat(512, 1024, 30208), one_shot=True)


def test_c4_poisson_n65536_properties():
    _check_posterior_properties(*_poisson(63488, 512), one_shot=False)
