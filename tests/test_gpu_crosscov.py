"""Process-vector cross-covariances and covariances of functionals (seam 1/3 of 8b):
``linfunctl(k, argnum)`` -> ``ProcessVectorCrossCovariance`` -> ``pv(x)`` / ``pv.evaluate_linop(x)``, ``linfunctl(pv)`` ->
``Covariance`` (src/linpde_gp/randprocs/crosscov/_pv_crosscov.py:14-161, crosscov/linfunctls/_evaluation.py:11-328,
src/linpde_gp/randvars/_covariance.py:13-230), against the numpy oracle on the same inputs.  Tolerance: 1e-12 of the
largest entry (the Gram gate of BASELINE.json's north_star)."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

TOL = 1e-12

MATERN_TP = {"scale": 2.5, "base": {"kind": "tensor_product", "factors": [
    {"kind": "matern", "nu": 2.5, "lengthscales": 0.6, "input_shape": []},
    {"kind": "matern", "nu": 3.5, "lengthscales": 0.9, "input_shape": []}]}}
EXPQUAD = {"scale": 1.7, "base": {"kind": "expquad", "lengthscales": [0.7, 1.1], "input_shape": [2]}}
NEG_LAP = [(-1.0, ("wl", [1.0, 1.0]))]
DD = [(1.0, ("dd", [0.3, -1.2]))]


def _close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-300), np.max(np.abs(a - b))


@pytest.mark.parametrize("kernel", [MATERN_TP, EXPQUAD], ids=["matern_tp", "expquad"])
@pytest.mark.parametrize("L1", [None, NEG_LAP, DD], ids=["id", "neglap", "dd"])
def test_pv_crosscov_of_point_evaluations(kernel, L1):
    """``(delta_X o L1)(k, argnum=1)``: x -> (k L1*)(x, X); shapes batch + (N,) and (N,) + batch, linop (M, N) / (N, M)."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.randprocs import crosscov
    from oracle import covfuncs as ocov

    rng = np.random.default_rng(0)
    X = rng.uniform(0, 1, (37, 2))
    Xt = rng.uniform(0, 1, (5, 4, 2))
    k = helpers.api_kernel(kernel)
    L = helpers.api_op(L1)
    fctl = lg.linfunctls._EvaluationFunctional((2,), (), X) if L is None else L.to_linfunctl(X)
    K_ref = ocov.matrix(kernel, None, L1, Xt.reshape(-1, 2), X)  # (20, 37)

    pv = fctl(k, argnum=1)
    assert isinstance(pv, crosscov.ProcessVectorCrossCovariance)
    assert pv.randproc_input_shape == (2,) and pv.randproc_output_shape == () and pv.randvar_shape == (37,)
    assert pv.randvar_size == 37 and not pv.reverse
    _close(pv(Xt), K_ref.reshape(5, 4, 37))
    op = pv.evaluate_linop(Xt)
    assert op.shape == (20, 37)
    _close(op.todense(), K_ref)

    pv_r = fctl(k, argnum=0)  # Cov(L f, f(.)) = the same numbers indexed (N,) + batch, since k is symmetric
    assert pv_r.reverse and pv_r.randvar_shape == (37,)
    K_ref_r = ocov.matrix(kernel, L1, None, X, Xt.reshape(-1, 2))  # (37, 20)
    _close(pv_r(Xt), K_ref_r.reshape(37, 5, 4))
    assert pv_r.evaluate_linop(Xt).shape == (37, 20)
    _close(pv_r.evaluate_linop(Xt).todense(), K_ref_r)

    with pytest.raises(ValueError):
        pv(np.zeros((3, 3)))  # trailing shape must equal the input shape of the process
    with pytest.raises(ValueError):
        fctl(k, argnum=2)


def test_pv_crosscov_arithmetic_and_operator_on_free_argument():
    """Scaled / summed cross-covariances (crosscov/_arithmetic.py) and ``L(pv)`` acting on the free argument
    (crosscov/linfuncops.py:18-87): ``L0(delta_X L1 (k, 1))(x) = (L0 k L1*)(x, X)``."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.randprocs import crosscov
    from oracle import covfuncs as ocov

    rng = np.random.default_rng(1)
    X, Xt = rng.uniform(0, 1, (21, 2)), rng.uniform(0, 1, (13, 2))
    k = helpers.api_kernel(MATERN_TP)
    lap, dd = helpers.api_op(NEG_LAP), helpers.api_op(DD)
    pv_lap = lap.to_linfunctl(X)(k, argnum=1)
    pv_id = lg.linfunctls._EvaluationFunctional((2,), (), X)(k, argnum=1)
    K_lap = ocov.matrix(MATERN_TP, None, NEG_LAP, Xt, X)
    K_id = ocov.matrix(MATERN_TP, None, None, Xt, X)

    s = 0.5 * pv_lap - 2.0 * pv_id
    assert isinstance(s, crosscov.SumProcessVectorCrossCovariance) and len(s.summands) == 2
    _close(s(Xt), 0.5 * K_lap - 2.0 * K_id)
    _close((-pv_lap)(Xt), -K_lap)
    assert isinstance(3.0 * (2.0 * pv_id), crosscov.ScaledProcessVectorCrossCovariance)
    _close((3.0 * (2.0 * pv_id)).evaluate_linop(Xt).todense(), 6.0 * K_id)

    # operators on the free argument
    _close(lap(pv_lap)(Xt), ocov.matrix(MATERN_TP, NEG_LAP, NEG_LAP, Xt, X))
    _close(dd(pv_id)(Xt), ocov.matrix(MATERN_TP, DD, None, Xt, X))
    _close(dd(s)(Xt), 0.5 * ocov.matrix(MATERN_TP, DD, NEG_LAP, Xt, X) - 2.0 * ocov.matrix(MATERN_TP, DD, None, Xt, X))
    # functionals of sums of functionals flatten into one cross-covariance
    f_sum = lap.to_linfunctl(X) + 0.25 * lg.linfunctls._EvaluationFunctional((2,), (), X)
    _close(f_sum(k, argnum=1)(Xt), K_lap + 0.25 * K_id)

    stacked = crosscov.StackedProcessVectorCrossCovariance((pv_lap, pv_id)).append(-pv_id)
    assert stacked.randvar_shape == (63,)
    _close(stacked(Xt), np.concatenate([K_lap, K_id, -K_id], axis=1))
    zero = lg.linfunctls._EvaluationFunctional((2,), (), X)(lg.randprocs.covfuncs.Zero((2,)), argnum=1)
    assert isinstance(zero, crosscov.Zero)
    _close(zero(Xt) + 1.0, np.ones((13, 21)))


@pytest.mark.parametrize("kernel", [MATERN_TP, EXPQUAD], ids=["matern_tp", "expquad"])
def test_covariance_of_two_functionals(kernel):
    """``L0_fctl(L1_fctl(k, argnum=1))`` -> ``Covariance`` with array / matrix / linop views
    (crosscov/linfunctls/_evaluation.py:11-18; randvars/_covariance.py): the Gram block of two observation batches."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200 import randvars
    from oracle import covfuncs as ocov

    rng = np.random.default_rng(2)
    X0, X1 = rng.uniform(0, 1, (3, 6, 2)), rng.uniform(0, 1, (17, 2))
    k = helpers.api_kernel(kernel)
    lap = helpers.api_op(NEG_LAP)
    f0 = lg.linfunctls._EvaluationFunctional((2,), (), X0)
    f1 = lap.to_linfunctl(X1)
    ref = ocov.matrix(kernel, None, NEG_LAP, X0.reshape(-1, 2), X1)  # (18, 17)

    cov = f0(f1(k, argnum=1))
    assert isinstance(cov, randvars.Covariance)
    assert cov.shape0 == (3, 6) and cov.shape1 == (17,) and cov.size0 == 18 and cov.size1 == 17
    _close(cov.matrix, ref)
    _close(cov.array, ref.reshape(3, 6, 17))
    _close(cov.linop.todense(), ref)
    _close(cov.T.matrix, ref.T)
    _close((2.0 * cov + cov).matrix, 3.0 * ref)
    assert cov.flatten0(np.zeros((3, 6))).shape == (18,) and cov.unflatten1(np.zeros(17)).shape == (17,)
    with pytest.raises(ValueError):
        cov.flatten0(np.zeros((6, 3)))

    # reversed order of application gives the transposed roles: Cov(L1 f, f(X0))
    cov_r = f0(f1(k, argnum=0))
    assert cov_r.shape0 == (17,) and cov_r.shape1 == (3, 6)
    _close(cov_r.matrix, ocov.matrix(kernel, NEG_LAP, None, X1, X0.reshape(-1, 2)))

    # both sides with operators: the L k L* Gram block, equal to the covariance function's own matrix
    cov_ll = lap.to_linfunctl(X0.reshape(-1, 2))(f1(k, argnum=1))
    _close(cov_ll.matrix, ocov.matrix(kernel, NEG_LAP, NEG_LAP, X0.reshape(-1, 2), X1))
    _close(cov_ll.matrix, lap(lap(k, argnum=1), argnum=0).matrix(X0.reshape(-1, 2), X1))


def test_crosscov_with_integral_functionals():
    """Lebesgue-integral functionals on a univariate Matern process (crosscov/linfunctls/integrals/): cross-covariance
    ``x -> int_a^b k(x, t) dt``, covariance with point evaluations and with another integral, and sums of both."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.randprocs import covfuncs
    from oracle import integrals as oint

    spec = {"scale": 1.5, "base": {"kind": "matern", "nu": 2.5, "lengthscales": 0.4, "input_shape": []}}
    k = 1.5 * covfuncs.Matern((), nu=2.5, lengthscales=0.4)
    dom, dom2 = (-0.5, 0.8), (0.1, 1.3)
    x = np.linspace(-1, 1, 29)
    I = lg.linfunctls.LebesgueIntegral(dom)
    I2 = lg.linfunctls.LebesgueIntegral(dom2)
    pv = I(k, argnum=1)
    assert pv.randvar_shape == () and pv.randvar_size == 1
    ref = oint.integral_crosscov(spec, dom, x)
    _close(pv(x), ref)
    assert pv.evaluate_linop(x).shape == (29, 1)
    _close(I(k, argnum=0).evaluate_linop(x).todense(), ref[None, :])

    ev = lg.linfunctls._EvaluationFunctional((), (), x)
    _close(ev(pv).matrix, ref[:, None])                     # Cov(f(x_i), int f)
    _close(I2(pv).matrix, [[oint.integral_integral(spec, dom2, dom)]])  # Cov(int_dom2 f, int_dom f)
    _close(I(ev(k, argnum=1)).matrix, ref[None, :])         # Cov(int f, f(x_j))

    y = np.array(0.3)  # a single point: output shape () like the integral's
    mixed = 2.0 * I - lg.linfunctls._EvaluationFunctional((), (), y)  # stationarity-type condition, one row
    pv_m = mixed(k, argnum=1)
    from oracle import covfuncs as ocov
    Ky = ocov.matrix(spec, None, None, x, y.reshape(1))[:, 0]
    _close(pv_m(x), 2.0 * ref - Ky)
    var_ref = (4.0 * oint.integral_integral(spec, dom, dom) - 4.0 * oint.integral_crosscov(spec, dom, y.reshape(1))[0]
               + ocov.matrix(spec, None, None, y.reshape(1), y.reshape(1))[0, 0])
    _close(mixed(pv_m).matrix, [[var_ref]], 1e-11)


def test_dirac_functional_conditioning_equals_evaluation_functional():
    """``DiracFunctional`` (linfunctls/_dirac.py) as an observation functional of a scalar process: same posterior as
    conditioning on ``X`` directly."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.randprocs import covfuncs

    rng = np.random.default_rng(4)
    X = rng.uniform(-1, 1, (4, 5))
    Y = np.sin(3 * X)
    prior = lg.GaussianProcess(lg.functions.Zero(()), 2.0 * covfuncs.Matern((), nu=2.5, lengthscales=0.5))
    dirac = lg.linfunctls.DiracFunctional((), (), X)
    assert dirac.output_shape == (4, 5) and dirac.X_batch_shape == (4, 5)
    np.testing.assert_allclose(dirac(lg.functions.Constant((), 1.5)), np.full((4, 5), 1.5))
    post_a = prior.condition_on_observations(Y, L=dirac)
    post_b = prior.condition_on_observations(Y.reshape(-1), X=X.reshape(-1))
    xs = np.linspace(-1, 1, 33)
    np.testing.assert_allclose(post_a.mean(xs), post_b.mean(xs), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(post_a.var(xs), post_b.var(xs), rtol=1e-10, atol=1e-12)
