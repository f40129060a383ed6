"""The oracle (numpy restatement, ``oracle/``) against the frozen outputs of the REAL reference
(``tests/golden/*.npz``, produced by ``oracle/make_golden.py``) and the reference's doctest known answers."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import covfuncs as ocf
from oracle import gp as ogp
from oracle import kron as okron
from tests.golden import cases as gcases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
_K = np.load(os.path.join(GOLDEN, "kernels.npz"))
SPECS = json.loads(bytes(_K["__specs__"]).decode())


@pytest.mark.parametrize("spec", SPECS, ids=[s["name"] for s in SPECS])
def test_kernel_matrix_matches_reference(spec):
    shape = gcases.kernel_input_shape(spec["kernel"])
    X = gcases.sobol_points(shape)
    K_ref, d_ref = _K[spec["name"] + "__K"], _K[spec["name"] + "__diag"]
    L0, L1 = gcases.spec_to_oracle_op(spec["L0"]), gcases.spec_to_oracle_op(spec["L1"])
    K = ocf.matrix(spec["kernel"], L0, L1, X[:32], X)
    d = ocf.diagonal(spec["kernel"], L0, L1, X[:32])
    scale = np.max(np.abs(K_ref))
    # the restatement follows the reference operation by operation: a few ulps at most
    assert np.max(np.abs(K - K_ref)) <= 4e-16 * scale
    assert np.max(np.abs(d - d_ref)) <= 4e-16 * scale


def test_generated_case_list_is_the_frozen_one():
    assert json.loads(json.dumps(gcases.build_cases())) == SPECS


def test_doctest_known_answers():
    # pn/randprocs/covfuncs/_matern.py:89-96
    k = {"scale": None, "base": {"kind": "matern", "input_shape": (), "nu": 2.5, "lengthscales": 0.1}}
    K = ocf.matrix(k, None, None, np.linspace(0, 1, 3))
    np.testing.assert_allclose(K[0], [1.0, 7.50933789e-04, 3.69569622e-08], rtol=1e-8)
    # pn/randprocs/covfuncs/_exponentiated_quadratic.py:45-52
    k = {"scale": None, "base": {"kind": "expquad", "input_shape": (), "lengthscales": 0.1}}
    K = ocf.matrix(k, None, None, np.linspace(0, 1, 3))
    np.testing.assert_allclose(K[0], [1.0, 3.72665317e-06, 1.92874985e-22], rtol=1e-8)


def test_matern_derivative_polynomial_table():
    """the table dumped from ``half_integer_matern_derivative_polynomial``."""
    from fractions import Fraction as F

    assert ocf.matern_derivative_polynomial(1, 4) == (F(-3), F(1))
    assert ocf.matern_derivative_polynomial(2, 2) == (F(-1, 3), F(-1, 3), F(1, 3))
    assert ocf.matern_derivative_polynomial(2, 4) == (F(1), F(-5, 3), F(1, 3))
    assert ocf.matern_derivative_polynomial(3, 4) == (F(1, 5), F(1, 5), F(-2, 5), F(1, 15))


def test_product_expquad_equals_ard_expquad():
    """tests/linpde_gp/randprocs/kernels/test_tensor_product.py:39-47."""
    X = gcases.sobol_points((3,))
    ls = [0.8, 1.1, 1.7]
    tp = {"scale": None, "base": {"kind": "tensor_product", "factors": [{"kind": "expquad", "input_shape": (), "lengthscales": l} for l in ls]}}
    ard = {"scale": None, "base": {"kind": "expquad", "input_shape": (3,), "lengthscales": ls}}
    np.testing.assert_allclose(ocf.matrix(tp, None, None, X), ocf.matrix(ard, None, None, X), rtol=1e-13, atol=1e-300)


GP_FILES = sorted(glob.glob(os.path.join(GOLDEN, "gp_*.npz")))


@pytest.mark.parametrize("path", GP_FILES, ids=[os.path.basename(p)[3:-4] for p in GP_FILES])
def test_gp_conditioning_matches_reference(path):
    g = np.load(path)
    problem = json.loads(bytes(g["problem"]).decode())
    res = ogp.solve(problem)
    np.testing.assert_allclose(res["gram"], g["gram"], rtol=0, atol=4e-16 * np.max(np.abs(g["gram"])))
    for key, tol in (("mean", 1e-9), ("var", 1e-8), ("cov", 1e-8)):
        sc = np.max(np.abs(g[key]))
        assert np.max(np.abs(res[key] - g[key])) <= tol * sc, key


def test_iterative_equals_batch_conditioning():
    """tests/linpde_gp/randprocs/test_posterior_gp.py:152-162 (iterative block conditioning == one shot)."""
    prob = ogp.golden_problems()["expquad_iterative"]
    it = ogp.solve(prob)
    X = np.concatenate([np.asarray(b["X"]) for b in prob["blocks"]])
    Y = np.concatenate([np.asarray(b["Y"]) for b in prob["blocks"]])
    nv = np.concatenate([np.full(len(b["Y"]), b["noise_var"] or 0.0) for b in prob["blocks"]])
    one = dict(prob, blocks=[{"X": X.tolist(), "Y": Y.tolist(), "L": None, "noise_var": nv.tolist()}])
    ob = ogp.solve(one)
    for key in ("mean", "var", "cov"):
        np.testing.assert_allclose(it[key], ob[key], rtol=1e-7, atol=1e-12)


def test_not_positive_definite_raises():
    from oracle import linalg as ola

    with pytest.raises(np.linalg.LinAlgError):
        ola.cholesky_lower(np.array([[1.0, 2.0], [2.0, 1.0]]))


# ---- tensor-grid (Kronecker) structure path -----------------------------------------------------------------------
_KR = np.load(os.path.join(GOLDEN, "kron.npz"))
KRON_SPECS = json.loads(bytes(_KR["__specs__"]).decode())


@pytest.mark.parametrize("spec", KRON_SPECS, ids=[s["name"] for s in KRON_SPECS])
def test_kronecker_linop_matches_reference(spec):
    """Frozen outputs of the reference's Kronecker linops on TensorProductGrids (densified + applied to vectors)
    against the restatement in oracle/kron.py -- and the Kronecker sum against the pairwise dense evaluation."""
    o0, o1 = gcases.spec_to_oracle_op(spec["L0"]), gcases.spec_to_oracle_op(spec["L1"])
    terms = okron.kronecker_terms(spec["kernel"], o0, o1, spec["factors0"], spec["factors1"])
    K_ref, V, KV_ref = _KR[spec["name"] + "__K"], _KR[spec["name"] + "__V"], _KR[spec["name"] + "__KV"]
    K = okron.dense(terms)
    sc = np.max(np.abs(K_ref))
    assert np.max(np.abs(K - K_ref)) <= 2e-15 * sc
    assert np.max(np.abs(okron.matvec(terms, V) - KV_ref)) <= 1e-13 * np.max(np.abs(KV_ref))
    d = len(spec["factors0"])
    g0 = okron.tensor_product_grid(*spec["factors0"]).reshape(-1, d)
    g1 = None if spec["factors1"] is None else okron.tensor_product_grid(*spec["factors1"]).reshape(-1, d)
    assert np.max(np.abs(ocf.matrix(spec["kernel"], o0, o1, g0, g1) - K_ref)) <= 2e-15 * sc


def test_generated_kron_case_list_is_the_frozen_one():
    assert json.loads(json.dumps(gcases.build_kron_cases())) == KRON_SPECS


# ---- multi-output processes with independent outputs ------------------------------------------------
MO_FILES = sorted(glob.glob(os.path.join(GOLDEN, "mo_*.npz")))


@pytest.mark.parametrize("path", MO_FILES, ids=[os.path.basename(p)[3:-4] for p in MO_FILES])
def test_multi_output_conditioning_matches_reference(path):
    """IndependentMultiOutputCovarianceFunction prior + ``D @ SelectOutput`` observations, run through the real
    reference by oracle/make_golden.py (experiments/0000_cpu_stationary_1d.ipynb cells 55-82 shape)."""
    from oracle import multi_output as omo

    g = np.load(path)
    problem = json.loads(bytes(g["problem"]).decode())
    res = omo.solve(problem)
    np.testing.assert_allclose(res["gram"], g["gram"], rtol=0, atol=4e-16 * np.max(np.abs(g["gram"])))
    for key, tol in (("mean", 1e-9), ("var", 1e-8), ("cov", 1e-8)):
        sc = np.max(np.abs(g[key]))
        assert np.max(np.abs(res[key] - g[key])) <= tol * sc, key
    assert len(MO_FILES) == 4


def test_multi_output_golden_problems_are_the_frozen_ones():
    from oracle import multi_output as omo

    for name, prob in omo.golden_problems().items():
        g = np.load(os.path.join(GOLDEN, f"mo_{name}.npz"))
        assert json.loads(json.dumps(prob)) == json.loads(bytes(g["problem"]).decode())


# ---- Lebesgue integrals of univariate half-integer Matern kernels (second half) ----------------------
_I = np.load(os.path.join(GOLDEN, "integrals.npz"))


def test_matern_lebesgue_integrals_match_reference():
    """The reference's own cases (tests/linpde_gp/randprocs/{crosscov,cov}/linfunctls/cases/cases_integral_matern.py):
    closed-form ``int_a^b k(x, t) dt`` at 10 points around the domain and ``int int k`` for three domain pairs, frozen
    outputs of the real ``UnivariateHalfIntegerMaternLebesgueIntegral``."""
    from oracle import integrals as oint

    for i, nu in enumerate(_I["nus"]):
        for j, ell in enumerate(_I["lengthscales"]):
            for d, (a, b) in enumerate(_I["domains"]):
                v = oint.matern_lebesgue_integral(int(nu - 0.5), ell, a, b, _I["X"][d])
                np.testing.assert_allclose(v, _I["Lk"][i, j, d], rtol=0, atol=1e-15)
            for d, (d0, d1) in enumerate(_I["domain_pairs"]):
                v = oint.matern_lebesgue_integral_lebesgue_integral(int(nu - 0.5), ell, tuple(d0), tuple(d1))
                assert abs(v - _I["LkL"][i, j, d]) <= 1e-15


def test_matern_lebesgue_integrals_match_quadrature():
    """The check the reference's tests make (test_Lk_kL.py / test_LkL.py): closed form == scipy.integrate.quad of the
    kernel (its generic fallback, _covfunc_lebesgue.py:45-51, 56-71)."""
    import scipy.integrate

    from oracle import covfuncs as ocf
    from oracle import integrals as oint

    for nu, ell in ((0.5, 0.8), (2.5, 1.1), (4.5, 2.1)):
        k = {"scale": None, "base": {"kind": "matern", "input_shape": [], "nu": nu, "lengthscales": ell}}
        kf = lambda s, t: float(ocf.evaluate(k, None, None, np.asarray(s), np.asarray(t)))  # noqa: E731
        (a, b), (c, d) = (-1.3, 0.0), (-0.2, 0.1)
        for x in (-2.0, -0.7, 0.0, 0.4):
            q = scipy.integrate.quad(lambda t: kf(x, t), a, b, points=[x] if a < x < b else None)[0]
            assert abs(oint.matern_lebesgue_integral(int(nu - 0.5), ell, a, b, x) - q) <= 1e-9
        q2 = scipy.integrate.dblquad(lambda t, s: kf(s, t), a, b, c, d, epsabs=1e-10)[0]
        assert abs(oint.matern_lebesgue_integral_lebesgue_integral(int(nu - 0.5), ell, (a, b), (c, d)) - q2) <= 1e-7


def test_oracle_crosscov_matches_reference_seam_goldens():
    """tests/golden/seams.npz (real reference: ``linfunctl(k, argnum=1)(x)``, ``evaluate_linop``, ``L0(L1(k, 1))``) against the
    oracle's ``matrix``: process-vector cross-covariances are the (M, N) matrices ``(k L*)(x, X)`` in the layouts
    batch + (N,) / (M, N) / shape0 + shape1 (crosscov/_pv_crosscov.py:60-161)."""
    g = np.load(os.path.join(GOLDEN, "seams.npz"))
    X, Xt, X0 = g["X"], g["Xt"], g["X0"]
    for kname, kspec in gcases.SEAM_KERNELS.items():
        for oname, ospec in gcases.SEAM_OPS.items():
            o_op = gcases.spec_to_oracle_op(ospec)
            K = ocf.matrix(kspec, None, o_op, Xt.reshape(-1, 2), X)
            sc = np.max(np.abs(K))
            assert np.max(np.abs(g[f"pv__{kname}__{oname}__argnum1"] - K.reshape(5, 4, 37))) <= 1e-14 * sc
            assert np.max(np.abs(g[f"pvlinop__{kname}__{oname}__argnum1"] - K)) <= 1e-14 * sc
            C = ocf.matrix(kspec, None, o_op, X0.reshape(-1, 2), X)
            assert np.max(np.abs(g[f"cov__{kname}__{oname}"] - C.reshape(3, 6, 37))) <= 1e-14 * np.max(np.abs(C))
            assert np.max(np.abs(g[f"covmat__{kname}__{oname}"] - C)) <= 1e-14 * np.max(np.abs(C))
    # the reference's bordered block factor is the Cholesky factor of the whole matrix (oracle/linalg.py restates it)
    for name in ("expquad3", "expquad_nested5", "matern200"):
        K = g[f"blk__{name}__K"]
        L = g[f"blk__{name}__chol"]
        assert np.max(np.abs(L @ L.T - K)) <= 1e-12 * np.max(np.abs(K))
        x = g[f"blk__{name}__schur_update"]
        rhs = np.concatenate([g[f"blk__{name}__u"], g[f"blk__{name}__v"]])
        assert np.max(np.abs(K @ x - rhs)) <= 1e-6 * np.max(np.abs(rhs))


def test_oracle_crosscov_matches_reference_linop_goldens():
    """tests/golden/seams_linop.npz (real reference: ``(B @ (A @ linfunctl))(k, argnum=1)(x)``, a
    ``LinOpProcessVectorCrossCovariance``, crosscov/_arithmetic.py:91-130) against the oracle's kernel matrix times
    ``(B A)^T``."""
    g = np.load(os.path.join(GOLDEN, "seams_linop.npz"))
    X, Xt, BA = g["X"], g["Xt"], g["B"] @ g["A"]
    n = 0
    for kname, kspec in gcases.SEAM_KERNELS.items():
        for oname, ospec in gcases.SEAM_OPS.items():
            K = ocf.matrix(kspec, None, gcases.spec_to_oracle_op(ospec), Xt.reshape(-1, 2), X) @ BA.T
            assert np.max(np.abs(g[f"pv__{kname}__{oname}"] - K.reshape(5, 4, 2))) <= 1e-13 * np.max(np.abs(K))
            n += 1
    assert n == len([k for k in g.files if k.startswith("pv__")])


def test_oracle_projections_match_reference_golden():
    """oracle/projections.py (quad / dblquad restatement + the Matern-3/2 closed form) against the real reference's
    outputs frozen in tests/golden/projections.npz; the cheap cases only (the quadrature cases take seconds each)."""
    import json

    from oracle import projections as oproj

    z = np.load(os.path.join(GOLDEN, "projections.npz"))
    for name in ("m32_ref", "m32_zb", "m12"):
        c = json.loads(bytes(z[f"{name}_spec"]).decode())
        ob = oproj.Basis(c["grid"], c["zero_boundary"])
        if c["kernel"]["nu"] == 1.5:
            val = oproj.crosscov_matern32(c["kernel"]["lengthscales"], ob, c["xs"], c["normalized"])
        else:
            val = oproj.crosscov_quad(c["kernel"], ob, c["xs"][:4], c["normalized"])
        ref = z[f"{name}_kPa"][: len(val)]
        assert np.max(np.abs(val - ref)) <= 1e-12 * np.max(np.abs(ref))
    # mass matrix of the basis: exact integrals of products of hat functions
    ob = oproj.Basis(np.array([0.0, 0.5, 1.5, 2.0]), False)
    M = oproj.mass_matrix(ob)
    assert np.allclose(M.sum(), 2.0) and np.allclose(M, M.T)
