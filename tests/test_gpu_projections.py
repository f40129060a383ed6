"""L2 projections onto piecewise-linear bases (SURVEY 8f item 4 tail;
src/linpde_gp/randprocs/crosscov/linfunctls/projections.py:18-170, linfunctls/projections/l2/_fem.py:14-95) through the
public API, against outputs of the REAL reference (tests/golden/projections.npz, oracle/make_golden.py) and against the
oracle's adaptive quadrature at tight tolerances.

Tolerances: the reference's closed form (Matern-3/2): 1e-12 of the largest entry; everything the reference computes with
scipy.integrate.quad / dblquad (default epsabs = epsrel = 1.49e-8): 1e-7; against the oracle's quadrature with
epsabs = 1e-13: 1e-11."""
import json
import os

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "projections.npz")
CASES = ["m32_ref", "m32_zb", "m52", "m12", "eq"]


def _golden():
    return np.load(GOLDEN)


def _spec(z, name):
    return json.loads(bytes(z[f"{name}_spec"]).decode())


def _proj(grid, zero_boundary=False, normalized=True):
    import linpde_gp_b200 as lg

    return lg.functions.bases.UnivariateLinearInterpolationBasis(np.asarray(grid), zero_boundary=zero_boundary).l2_projection(
        normalized=normalized)


def _rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b))


@pytest.mark.parametrize("name", CASES)
def test_projection_crosscov_matches_reference(name):
    """``proj(k, argnum=1)(xs)`` (batch + (m,)) and ``argnum=0`` ((m,) + batch) against the reference's values."""
    z = _golden()
    c = _spec(z, name)
    k = helpers.api_base(c["kernel"])
    proj = _proj(c["grid"], c["zero_boundary"], c["normalized"])
    xs = np.asarray(c["xs"])
    kPa = proj(k, argnum=1)
    from linpde_gp_b200.randprocs import crosscov

    assert isinstance(kPa, crosscov.CovarianceFunction_L2Projection_UnivariateLinearInterpolationBasis)
    assert isinstance(kPa, crosscov.Matern32_L2Projection_UnivariateLinearInterpolationBasis) == (c["kernel"].get("nu") == 1.5)
    assert kPa.projection is proj and kPa.covfunc is k
    assert kPa.randvar_shape == (len(proj.basis),) and not kPa.reverse
    val = kPa(xs)
    ref = z[f"{name}_kPa"]
    assert val.shape == ref.shape
    closed_form = c["kernel"]["kind"] == "matern" and c["kernel"]["nu"] == 1.5
    assert _rel(val, ref) <= (1e-12 if closed_form else 1e-7), _rel(val, ref)
    rev = proj(k, argnum=0)(xs)
    assert rev.shape == (len(proj.basis),) + xs.shape and np.array_equal(rev, np.moveaxis(val, -1, 0))
    # 2-d batch of points
    val2 = kPa(xs[:12].reshape(3, 4))
    assert val2.shape == (3, 4, len(proj.basis)) and np.array_equal(val2.reshape(12, -1), val[:12])


@pytest.mark.parametrize("nu,ell", [(0.5, 0.5), (1.5, 0.2), (2.5, 0.7), (3.5, 1.3)])
@pytest.mark.parametrize("zero_boundary", [False, True])
def test_closed_form_hat_integrals_vs_tight_quadrature(nu, ell, zero_boundary):
    """``lpgp_matern_hat_integral`` for every half-integer order against the oracle's adaptive quadrature at
    epsabs = 1e-13 (integrand split at the kink x), un-normalised and normalised, non-uniform nodes, points inside,
    on and far outside the grid."""
    import scipy.integrate

    from oracle import covfuncs as ocf
    from oracle import projections as oproj

    grid = np.array([-0.4, -0.1, 0.0, 0.25, 0.6, 0.7, 1.2])
    xs = np.array([-3.0, -0.4, -0.25, -0.1, 0.0, 0.1, 0.25, 0.65, 0.7, 1.0, 1.2, 1.5, 6.0])
    kernel = {"scale": None, "base": {"kind": "matern", "input_shape": [], "nu": nu, "lengthscales": ell}}
    ob = oproj.Basis(grid, zero_boundary)
    kf = lambda x, t: float(ocf.matrix(kernel, None, None, np.array([x]), np.array([t]))[0, 0])  # noqa: E731
    ref = np.zeros((len(xs), len(ob)))
    for i, x in enumerate(xs):
        for j in range(len(ob)):
            a, b = ob.support_bounds(j)
            pts = sorted({a, b, float(ob.x_i[j]), min(max(x, a), b)})
            ref[i, j] = sum(scipy.integrate.quad(lambda t, j=j, x=x: float(ob.eval_elem(j, t)) * kf(x, t), lo, hi,
                                                 epsabs=1e-13, epsrel=1e-13)[0] for lo, hi in zip(pts[:-1], pts[1:]))
    k = helpers.api_base(kernel["base"])
    val = _proj(grid, zero_boundary, normalized=False)(k, argnum=1)(xs)
    assert np.max(np.abs(val - ref)) <= 1e-11 * np.max(np.abs(ref)), np.max(np.abs(val - ref))
    valn = _proj(grid, zero_boundary, normalized=True)(k, argnum=1)(xs)
    refn = oproj.normalize(ob, ref, -1, True)
    assert np.max(np.abs(valn - refn)) <= 1e-10 * np.max(np.abs(refn))


@pytest.mark.parametrize("name", ["m32", "eq"])
def test_covariance_of_two_projections_matches_reference(name):
    """``proj(proj(k, argnum=1))`` -> Covariance (m x m): dblquad in the reference, closed form / Gauss-Legendre here."""
    z = _golden()
    kernel = json.loads(bytes(z[f"PkP_{name}_kernel"]).decode())
    proj = _proj(z[f"PkP_{name}_grid"])
    C = proj(proj(helpers.api_base(kernel), argnum=1))
    arr = np.asarray(C.array)
    ref = z[f"PkP_{name}"]
    assert arr.shape == ref.shape
    assert _rel(arr, ref) <= 1e-7, _rel(arr, ref)
    assert _rel(arr, arr.T) <= 1e-12  # symmetric to quadrature accuracy


def test_projection_of_functions():
    """``proj(f)`` for constants (closed form, _fem.py:83-95) and general functions (quadrature, :64-81): the projection
    of a piecewise-linear function on the same nodes is its vector of nodal values."""
    import linpde_gp_b200 as lg

    grid = np.array([0.0, 0.2, 0.5, 0.6, 1.0])
    proj = _proj(grid)
    c = proj(lg.functions.Constant(input_shape=(), value=2.5))
    assert np.allclose(c, 2.5, rtol=0, atol=1e-12)
    f = lg.functions.LambdaFunction(lambda x: 3.0 * x - 1.0, input_shape=())
    assert np.allclose(proj(f), 3.0 * grid - 1.0, rtol=0, atol=1e-12)
    un = _proj(grid, normalized=False)(lg.functions.Constant(input_shape=(), value=1.0))
    assert np.allclose(un, [0.1, 0.25, 0.2, 0.25, 0.2], rtol=0, atol=1e-14)


def test_conditioning_on_projection_then_point_observations_matches_reference():
    """``prior.condition_on_observations(Y, L=proj)`` followed by point observations, against the real reference
    (whose Gram blocks come from dblquad / quad: 1e-7); the Gram matrix block of the projection is P k P*."""
    import linpde_gp_b200 as lg

    z = _golden()
    kernel = json.loads(bytes(z["gp_kernel"]).decode())
    prior = lg.GaussianProcess(lg.functions.Constant(input_shape=(), value=0.3), 2.0 * helpers.api_base(kernel))
    proj = _proj(z["gp_grid"])
    post1 = prior.condition_on_observations(z["gp_Yp"], L=proj)
    Xt = z["gp_Xt"]
    sc = 2.0
    assert np.max(np.abs(post1.mean(Xt) - z["gp_mean1"])) <= 1e-6 * np.max(np.abs(z["gp_mean1"]))
    assert np.max(np.abs(post1.var(Xt) - z["gp_var1"])) <= 1e-6 * sc
    post2 = post1.condition_on_observations(z["gp_Yo"], X=z["gp_Xo"])
    assert np.max(np.abs(np.asarray(post2.gram.todense()) - z["gp_gram2"])) <= 1e-7 * np.max(np.abs(z["gp_gram2"]))
    assert np.max(np.abs(post2.mean(Xt) - z["gp_mean2"])) <= 1e-6 * np.max(np.abs(z["gp_mean2"]))
    assert np.max(np.abs(post2.var(Xt) - z["gp_var2"])) <= 1e-6 * sc
    # the other order (points first, projection second) is not available in the reference; the bordered factor makes
    # it the same posterior here
    post3 = prior.condition_on_observations(z["gp_Yo"], X=z["gp_Xo"]).condition_on_observations(z["gp_Yp"], L=proj)
    assert np.max(np.abs(post3.mean(Xt) - post2.mean(Xt))) <= 1e-9 * np.max(np.abs(z["gp_mean2"]))
    assert np.max(np.abs(post3.var(Xt) - post2.var(Xt))) <= 1e-9 * sc
    # posterior covariance block is consistent with the pointwise variance
    C = np.asarray(post2.cov.linop(Xt[:7]).todense())
    assert np.max(np.abs(np.diag(C) - post2.var(Xt[:7]))) <= 1e-10 * sc


def test_projection_observation_of_an_expquad_process_interpolates():
    """Smooth kernel (quadrature path): after conditioning on ``P f = y`` the posterior mean's projection is ``y`` and the
    projected posterior has no variance left."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.randprocs import covfuncs

    grid = np.linspace(0.0, 1.0, 9)
    proj = _proj(grid, zero_boundary=True)
    prior = lg.GaussianProcess(lg.functions.Zero(input_shape=()), 1.5 * covfuncs.ExpQuad((), lengthscales=0.35))
    y = np.cos(3.0 * grid[1:-1])
    post = prior.condition_on_observations(y, L=proj)
    nodes, W = proj.basis.gauss_legendre(24)
    Pm = proj.normalizer(W @ post.mean(nodes), axis=-1)
    assert np.max(np.abs(Pm - y)) <= 1e-8
