"""Host-side logic of the reference-style API that needs no GPU: symbolic operator algebra, dispatch types,
argument validation / error behaviour, the C-ABI library's exported symbols."""
import ctypes
import os
import re

import numpy as np
import pytest

import linpde_gp_b200 as lg
from linpde_gp_b200.linfuncops import diffops
from linpde_gp_b200.randprocs import covfuncs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "lpgp.h")).read()
    declared = set(re.findall(r"\b(lpgp_[a-z0-9_]+)\s*\(", header))
    lib = ctypes.CDLL(os.path.join(ROOT, "linpde_gp_b200", "lib", "liblpgp.so"))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), name
    assert set(lg._lib.EXPORTED) == declared
    assert lib.lpgp_version() >= 100
    lib.lpgp_build_arch.restype = ctypes.c_char_p
    assert lib.lpgp_build_arch() == b"sm_100a"


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(lg._lib.KernelDesc) == 4 + 3 * 16 + 4 + 4 * 8 + 8 + 512 * 8
    assert ctypes.sizeof(lg._lib.Factor) == 8 + 8 + 8 + 8 + 4 + 4 + 65 * 8
    assert ctypes.sizeof(lg._lib.ObsBlock) == 32


def test_coefficients_algebra():
    """tests/linpde_gp/linfuncops/diffops/test_coefficients.py, test_laplacian.py"""
    lap = diffops.Laplacian((2,))
    c = lap.coefficients
    assert c.num_entries == 2 and not c.has_mixed
    assert c[()][diffops.MultiIndex((2, 0))] == 1.0 and c[()][diffops.MultiIndex((0, 2))] == 1.0
    neg = (-2.0 * lap).coefficients
    assert neg[()][diffops.MultiIndex((0, 2))] == -2.0
    s = c + diffops.DirectionalDerivative([1.0, 0.0]).coefficients
    assert s.num_entries == 3
    with pytest.raises(ValueError):
        c + diffops.Laplacian((3,)).coefficients
    with pytest.raises(ValueError):
        diffops.MultiIndex((-1, 0))
    wl = diffops.WeightedLaplacian([0.0, 3.0])
    assert wl.coefficients.num_entries == 1  # zero weights dropped (_laplacian.py:31-41)
    assert diffops.SpatialLaplacian((3,)).weights.tolist() == [0.0, 1.0, 1.0]
    with pytest.raises(ValueError):
        diffops.SpatialLaplacian((1,))
    heat = diffops.HeatOperator((2,), alpha=0.1)
    assert heat._terms() == {(1, 0): 1.0, (0, 2): -0.1}
    with pytest.raises(ValueError):
        diffops.HeatOperator((2, 2))
    assert diffops.TimeDerivative((3,)).multi_index.as_tuple() == (1, 0, 0)
    assert (3.0 * (2.0 * lap)).scalar == 6.0


def test_dispatch_types_follow_the_reference_registry():
    """test_diffops.py::test_L0kL1_expected_type / SURVEY Appendix A."""
    eq = covfuncs.ExpQuad((3,))
    wl, dd = diffops.WeightedLaplacian(np.ones(3)), diffops.DirectionalDerivative(np.ones(3))
    assert type(wl(eq, argnum=1)) is covfuncs.ExpQuad_Identity_WeightedLaplacian
    assert type(wl(wl(eq, argnum=1), argnum=0)) is covfuncs.ExpQuad_WeightedLaplacian_WeightedLaplacian
    assert type(dd(eq, argnum=0)) is covfuncs.ExpQuad_Identity_DirectionalDerivative
    assert type(dd(dd(eq, argnum=1), argnum=0)) is covfuncs.ExpQuad_DirectionalDerivative_DirectionalDerivative
    assert type(dd(wl(eq, argnum=1), argnum=0)) is covfuncs.ExpQuad_DirectionalDerivative_WeightedLaplacian
    m = covfuncs.Matern((), nu=2.5)
    w1, d1 = diffops.WeightedLaplacian(2.0), diffops.DirectionalDerivative(1.5)
    assert type(w1(m, argnum=1)) is covfuncs.UnivariateHalfIntegerMatern_Identity_WeightedLaplacian
    assert type(w1(w1(m, argnum=1), argnum=0)) is covfuncs.UnivariateHalfIntegerMatern_WeightedLaplacian_WeightedLaplacian
    assert type(d1(m, argnum=1)) is covfuncs.HalfIntegerMatern_Identity_DirectionalDerivative
    assert type(d1(w1(m, argnum=1), argnum=0)) is covfuncs.UnivariateHalfIntegerMatern_DirectionalDerivative_WeightedLaplacian
    tp = covfuncs.TensorProduct(covfuncs.Matern((), nu=1.5), covfuncs.Matern((), nu=2.5))
    heat = diffops.HeatOperator((2,), 0.1)
    assert type(heat(heat(tp, argnum=1), argnum=0)) is covfuncs.TensorProduct_LinDiffOp_LinDiffOp
    scaled = 4.0 * tp
    out = (-1.0 * diffops.Laplacian((2,)))(scaled, argnum=1)
    assert type(out) is covfuncs.ScaledCovarianceFunction and float(out.scalar) == 4.0
    assert type(out.covfunc) is covfuncs.TensorProduct_LinDiffOp_LinDiffOp  # the -1 is folded into the operator terms
    assert out.covfunc.L1._terms() == {(2, 0): -1.0, (0, 2): -1.0}
    # isotropic multi-d Matern: first-order operators only (diffops/_registry.py:142-190)
    m3, d3, e3 = covfuncs.Matern((3,), nu=2.5), diffops.DirectionalDerivative(np.ones(3)), diffops.DirectionalDerivative(np.arange(3.0))
    assert type(d3(m3, argnum=1)) is covfuncs.HalfIntegerMatern_Identity_DirectionalDerivative
    assert type(e3(d3(m3, argnum=1), argnum=0)) is covfuncs.HalfIntegerMatern_DirectionalDerivative_DirectionalDerivative
    assert type(e3(d3(m3, argnum=0), argnum=1)) is covfuncs.HalfIntegerMatern_DirectionalDerivative_DirectionalDerivative
    assert type(d1(d1(m, argnum=1), argnum=0)) is covfuncs.UnivariateHalfIntegerMatern_DirectionalDerivative_DirectionalDerivative
    # isotropic multi-d Matern x Laplacian: no closed form (reference falls back to jax, _registry.py:270-280)
    with pytest.raises(NotImplementedError):
        diffops.Laplacian((2,))(covfuncs.Matern((2,), nu=2.5), argnum=1)
    with pytest.raises(ValueError):
        diffops.Laplacian((3,))(tp, argnum=0)
    with pytest.raises(ValueError):
        tp_bad = covfuncs.TensorProduct(covfuncs.Matern((2,), nu=1.5))


def test_covfunc_argument_validation():
    with pytest.raises(ValueError):
        covfuncs.Matern((), nu=-1.0)
    with pytest.raises(ValueError):
        covfuncs.ExpQuad((2,), lengthscales=[1.0, -1.0])
    k = covfuncs.ExpQuad((2,))
    with pytest.raises(ValueError):
        k(np.zeros((4, 3)), None)
    with pytest.raises(ValueError):
        k(np.zeros((4, 2)), np.zeros((3, 2)))  # batch shapes do not broadcast
    with pytest.raises(ValueError):
        k.linop(np.zeros((4, 3)))


def test_condition_on_observations_error_behaviour():
    """_conditional.py:317-387: the same exception types as the reference."""
    prior = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), covfuncs.ExpQuad((2,)))
    X = np.zeros((5, 2))
    L = diffops.Laplacian((2,))
    with pytest.raises(ValueError):
        prior.condition_on_observations(np.zeros(5))  # X and L both omitted
    with pytest.raises(ValueError):
        prior.condition_on_observations(np.zeros(5), L=L)  # operator without X
    with pytest.raises(TypeError):
        prior.condition_on_observations(np.zeros(5), X=X, L=L.to_linfunctl(X))  # functional with X
    with pytest.raises(TypeError):
        prior.condition_on_observations(np.zeros(5), X=X, L="laplacian")
    with pytest.raises(ValueError):
        prior.condition_on_observations(np.zeros(4), X=X, L=L)  # Y shape mismatch
    with pytest.raises(ValueError):
        prior.condition_on_observations(np.zeros(5), X=X, b=lg.randvars.Normal(np.zeros(4), np.eye(4)))
    with pytest.raises(ValueError):
        lg.GaussianProcess(lg.functions.Zero(input_shape=(3,)), covfuncs.ExpQuad((2,)))
    with pytest.raises(TypeError):
        lg.GaussianProcess(lambda x: x, covfuncs.ExpQuad((2,)))


def test_product_path_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    k = covfuncs.ExpQuad((2,))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        k.matrix(np.zeros((3, 2)))
    # the product never imports the oracle
    import sys

    src = []
    for root, _, files in os.walk(os.path.join(ROOT, "linpde_gp_b200")):
        src += [os.path.join(root, f) for f in files if f.endswith((".py", ".cu", ".cuh"))]
    for path in src:
        text = open(path).read()
        assert "import oracle" not in text and "from oracle" not in text, path


def test_tensor_product_grid_host_semantics():
    """TensorProductGrid (src/linpde_gp/randprocs/covfuncs/_tensor_product.py:133-152): meshgrid "ij" stacked on the last
    axis; the C-order flattening enumerates points in Kronecker order; views lose the factorisation."""
    from linpde_gp_b200.randprocs import covfuncs

    xs, ys = np.array([0.0, 0.5, 1.0]), np.array([-1.0, 1.0])
    g = covfuncs.TensorProductGrid(xs, ys)
    assert g.shape == (3, 2, 2) and isinstance(g, np.ndarray)
    assert len(g.factors) == 2 and np.array_equal(g.factors[0], xs)
    flat = np.asarray(g).reshape(-1, 2)
    np.testing.assert_array_equal(flat[:, 0], np.kron(xs, np.ones(2)))
    np.testing.assert_array_equal(flat[:, 1], np.kron(np.ones(3), ys))
    assert covfuncs._grid_factors(g) is g.factors
    assert covfuncs._grid_factors(g[1:]) is None and covfuncs._grid_factors(g + 1.0) is None
    assert covfuncs._grid_factors(flat) is None
    assert covfuncs._grid_factors(covfuncs.TensorProductGrid(xs, ys, indexing="xy")) is None
    with pytest.raises(ValueError):
        covfuncs.TensorProductGrid(np.zeros((2, 2)))


def test_multi_output_dispatch_follows_the_reference_registry():
    """covfuncs/linfuncops/_registry.py:34-48, 82-120 and linfuncops/_arithmetic.py:112-144 (host-side, symbolic)."""
    from linpde_gp_b200 import functions, linfuncops
    from linpde_gp_b200.randprocs._conditional import _descs

    ku = 9.0 * covfuncs.Matern((), nu=2.5, lengthscales=0.75)
    kv = 0.81 * covfuncs.Matern((), nu=0.5)
    k = covfuncs.IndependentMultiOutputCovarianceFunction(ku, kv, kv)
    assert k.output_shape_0 == (3,) and k.output_shape_1 == (3,) and k.input_shape == ()
    su, sv, sa = (linfuncops.SelectOutput(((), (3,)), idx=i) for i in range(3))
    assert su.output_shapes == ((), ()) and su.input_shapes == ((), (3,))
    D = -0.5 * diffops.Laplacian(())
    # SelectOutput on the block-diagonal kernel -> a stack with zeros off the selected output
    st = su(k, argnum=1)
    assert type(st) is covfuncs.StackCovarianceFunction and st.output_idx == 0 and st.output_shape_0 == (3,)
    assert st.covfuncs[0] is ku and all(type(c) is covfuncs.Zero for c in st.covfuncs[1:])
    assert su(k, argnum=0).output_idx == 1
    # selecting the stacked argument picks the entry; selecting twice recovers the output's own kernel
    assert su(st, argnum=0) is ku and type(sv(st, argnum=0)) is covfuncs.Zero
    # composition acts right to left; sums / scalars distribute (Sum of stacks stays a stack)
    L = D @ su - sv
    assert type(D @ su) is linfuncops.CompositeLinearFunctionOperator
    assert type(D @ (su + su)) is linfuncops.SumLinearFunctionOperator
    kL = L(k, argnum=1)
    assert type(kL) is covfuncs.StackCovarianceFunction and kL.output_idx == 0
    assert type(kL.covfuncs[2]) is covfuncs.Zero
    inner = kL.covfuncs[0]
    assert float(inner.scalar) == 9.0
    assert type(inner.covfunc) is covfuncs.UnivariateHalfIntegerMatern_Identity_WeightedLaplacian
    LkL = L(kL, argnum=0)
    assert type(LkL) is covfuncs.SumCovarianceFunction and len(LkL.summands) == 2 and LkL.output_shape_0 == ()
    assert type(LkL.summands[0].covfunc) is covfuncs.UnivariateHalfIntegerMatern_WeightedLaplacian_WeightedLaplacian
    assert len(_descs(LkL)) == 2 and _descs(sa(kL, argnum=0)) == []
    # an operator on the stacked argument needs a SelectOutput first (validate_covfunc_transformation: ValueError)
    with pytest.raises(ValueError):
        D(kL, argnum=0)
    with pytest.raises(ValueError):
        D(k, argnum=0)  # codomain shapes differ
    # functions: SelectOutput / linear combinations of a stacked mean
    mean = functions.StackedFunction(functions.Constant((), 57.0), functions.Constant((), 0.3), functions.Constant((), -0.1))
    assert mean.output_shape == (3,) and mean(np.zeros(4)).shape == (4, 3)
    np.testing.assert_allclose(L(mean)(np.zeros(2)), [-0.3, -0.3])
    np.testing.assert_allclose((2.0 * su - sa)(mean)(np.zeros(1)), [114.1])
    # processes
    prior = lg.GaussianProcess(mean, k)
    assert prior.output_shape == (3,)
    pu = su(prior)
    assert pu.output_shape == () and pu.cov is ku
    with pytest.raises(ValueError):
        lg.GaussianProcess(functions.Constant((), 1.0), k)
    with pytest.raises(NotImplementedError):
        prior.condition_on_observations(np.zeros((3, 4)), X=np.zeros(4))  # vector-valued observation


def test_composite_linear_functional_keywords_follow_the_reference():
    """src/linpde_gp/linfunctls/_arithmetic.py:92-174, _linfunctl.py:103-129: ``linop`` is the MATRIX applied last,
    ``linfuncop`` the function operator applied first; ``A @ linfunctl`` / ``linfunctl @ L`` build the composite."""
    from linpde_gp_b200 import functions, linfunctls, linops
    from linpde_gp_b200.linfuncops import diffops

    X = np.linspace(0.0, 1.0, 5)
    ev = linfunctls._EvaluationFunctional(input_domain_shape=(), input_codomain_shape=(), X=X)
    D = diffops.Derivative(1)
    A = np.arange(15.0).reshape(3, 5)
    f = functions.Polynomial([1.0, 2.0, 3.0])
    plain = linfunctls.CompositeLinearFunctional(linop=None, linfunctl=ev, linfuncop=D)  # the reference's spelling
    assert plain.linop is None and plain.linfuncop is D and plain.linfunctl is ev and plain.output_shape == (5,)
    assert len(plain._atoms()) == 1 and plain._atoms()[0][2] is D
    assert type(ev @ D) is linfunctls.CompositeLinearFunctional and (ev @ D).linfuncop is D
    comp = A @ ev
    assert type(comp) is linfunctls.CompositeLinearFunctional
    assert comp.output_shape == (3,) and comp.input_shapes == ((), ()) and comp.linfuncop is None
    np.testing.assert_allclose(comp.linop, A)
    np.testing.assert_allclose(comp(f), A @ f(X), rtol=1e-14)
    comp2 = np.ones((2, 3)) @ (comp @ D)  # (B A) @ ev @ D
    assert comp2.output_shape == (2,) and comp2.linfuncop is D
    np.testing.assert_allclose(comp2.linop, np.ones((2, 3)) @ A)
    np.testing.assert_allclose(comp2(f), np.ones((2, 3)) @ A @ D(f)(X), rtol=1e-14)

    class _HostOp(linops.LinearOperator):  # a LinearOperator is accepted like an array (densified once)
        def __init__(self):
            super().__init__(A.shape)

        def todense(self, cache=True):
            return A

    np.testing.assert_allclose((_HostOp() @ ev).linop, A)
    with pytest.raises(NotImplementedError):  # conditioning on matrix @ functional is not lowered to the device
        comp._atoms()
    with pytest.raises(ValueError):
        linfunctls.CompositeLinearFunctional(linop=np.ones((3, 4)), linfunctl=ev, linfuncop=None)
    with pytest.raises(ValueError):  # the matrix needs a 1-D functional output
        np.ones((2, 1)) @ linfunctls.LebesgueIntegral((0.0, 1.0))
    legacy = linfunctls.CompositeLinearFunctional(linop=D, linfunctl=ev)  # round-1 spelling still understood
    assert legacy.linop is None and legacy.linfuncop is D
    with pytest.raises(TypeError):
        linfunctls.CompositeLinearFunctional(linop=D, linfunctl=ev, linfuncop=D)


def test_linear_functional_arithmetic_and_atoms():
    """src/linpde_gp/linfunctls/_linfunctl.py:74-129, _arithmetic.py:12-174, _integrals.py:13-62 (host-side, symbolic):
    the stationarity functional of experiments/0000_cpu_stationary_1d.ipynb cell 65."""
    from linpde_gp_b200 import _lowering, functions, linfuncops, linfunctls
    from oracle import integrals as oint

    sel = [linfuncops.SelectOutput(((), (3,)), idx=j) for j in range(3)]
    integral = linfunctls.LebesgueIntegral((0.0, 1.5))
    assert integral.domain == (0.0, 1.5) and integral.output_shape == () and integral.input_shapes == ((), ())
    part = 0.4 * integral @ sel[1]
    assert type(0.4 * integral) is linfunctls.ScaledLinearFunctional
    assert type(part) is linfunctls.CompositeLinearFunctional and part.linfuncop is sel[1]
    assert part.input_shapes == ((), (3,)) and part.output_shape == ()
    L = part + 0.4 * (sel[2].to_linfunctl(1.5) + sel[2].to_linfunctl(0.0))
    assert type(L) is linfunctls.SumLinearFunctional and L.output_shape == ()
    atoms = L._atoms()
    assert [(c, kind) for c, kind, _, _ in atoms] == [(0.4, "int"), (0.4, "pts"), (0.4, "pts")]
    assert atoms[0][2] is sel[1] and atoms[0][3] == (0.0, 1.5) and float(atoms[1][3]) == 1.5
    assert type(2.0 * (0.4 * integral)) is linfunctls.ScaledLinearFunctional and float((2.0 * (0.4 * integral)).scalar) == 0.8
    assert [c for c, *_ in (-L)._atoms()] == [-0.4, -0.4, -0.4]
    assert [c for c, *_ in (part - part)._atoms()] == [0.4, -0.4]
    # functions: integral of a constant = value * volume; generic functions by quadrature (like the reference)
    mean = functions.StackedFunction(functions.Constant((), 57.0), functions.Constant((), 0.3), functions.Constant((), -0.1))
    np.testing.assert_allclose(L(mean), 0.4 * 0.3 * 1.5 + 0.4 * 2 * (-0.1), rtol=1e-15)
    np.testing.assert_allclose(integral(functions.LambdaFunction(lambda x: np.sin(x), (), ())), 1 - np.cos(1.5), rtol=1e-12)
    # a point-evaluation functional still exposes (operator, points); composite functionals do not
    op, X = sel[0].to_linfunctl(np.linspace(0, 1, 4))._as_observation()
    assert op is sel[0] and X.shape == (4,)
    with pytest.raises(NotImplementedError):
        L._as_observation()
    # argument / shape errors
    with pytest.raises(ValueError):
        linfunctls.SumLinearFunctional(integral, sel[0].to_linfunctl(np.zeros(3)))  # output shapes differ
    with pytest.raises(ValueError):
        integral @ linfuncops.SelectOutput(((2,), (3,)), idx=0)  # domain shapes differ
    with pytest.raises(TypeError):
        linfunctls.LebesgueIntegral(3.0)
    with pytest.raises(NotImplementedError):
        linfunctls.LebesgueIntegral([[0.0, 1.0], [0.0, 1.0]])
    # exact antiderivative polynomials of the device descriptor == the oracle's restatement of the reference classes
    for p in range(6):
        p1, p2 = _lowering.matern_antiderivative_polys(p)
        assert p1 == oint.antiderivative_polynomial(p) and p2 == oint.second_antiderivative_polynomial(p)
    dsc = _lowering.matern_integral_desc(2.5, 0.5)
    assert dsc.ncoef == 3 and dsc.scale == np.sqrt(5.0) / 0.5 and dsc.poly1[0] == float(oint.antiderivative_polynomial(2)[0])
    with pytest.raises(NotImplementedError):
        _lowering.matern_integral_desc(1.2, 1.0)


def test_integral_dispatch_needs_closed_form_kernels():
    """covfuncs/linfunctls/_registry.py:157-193: Scaled / Sum / univariate half-integer Matern have closed forms; the
    reference's scipy quadrature fallback for everything else is not provided (NotImplementedError, no CPU path)."""
    from linpde_gp_b200.randprocs._conditional import _integral_terms

    m1 = covfuncs.Matern((), nu=1.5, lengthscales=0.8)
    m2 = covfuncs.Matern((), nu=0.5, lengthscales=2.0)
    assert _integral_terms(3.0 * m1 + 0.5 * (2.0 * m2)) == [(3.0, 1.5, 0.8), (1.0, 0.5, 2.0)]
    assert _integral_terms(covfuncs.Zero(())) == []
    with pytest.raises(NotImplementedError):
        _integral_terms(covfuncs.ExpQuad((), lengthscales=1.0))
    with pytest.raises(NotImplementedError):
        _integral_terms(diffops.Laplacian(())(m1, argnum=1))
    with pytest.raises(NotImplementedError):
        _integral_terms(covfuncs.Matern((), nu=1.2))


def test_covariance_containers_host_semantics():
    """``randvars.ArrayCovariance`` / ``Covariance`` shape bookkeeping (src/linpde_gp/randvars/_covariance.py:13-194): array
    <-> matrix views over C-order flattened variables, flatten / unflatten with shape errors."""
    from linpde_gp_b200 import randvars

    arr = np.arange(2 * 3 * 4, dtype=float).reshape(2, 3, 4)
    cov = randvars.ArrayCovariance(arr, shape0=(2, 3), shape1=(4,))
    assert cov.shape0 == (2, 3) and cov.shape1 == (4,) and cov.ndim0 == 2 and cov.ndim1 == 1
    assert cov.size0 == 6 and cov.size1 == 4
    np.testing.assert_array_equal(cov.array, arr)
    np.testing.assert_array_equal(cov.matrix, arr.reshape(6, 4))
    assert cov.flatten0(np.zeros((2, 3))).shape == (6,) and cov.flatten1(np.zeros(4)).shape == (4,)
    assert cov.unflatten0(np.zeros(6)).shape == (2, 3) and cov.unflatten1(np.zeros(4)).shape == (4,)
    with pytest.raises(ValueError):
        cov.flatten0(np.zeros((3, 2)))
    with pytest.raises(ValueError):
        cov.flatten1(np.zeros(5))
    with pytest.raises(ValueError):
        randvars.ArrayCovariance(arr, shape0=(2, 3), shape1=(5,))
    s = randvars.ArrayCovariance.from_scalar(2.5)
    assert s.shape0 == () and s.shape1 == () and s.size0 == 1 and s.matrix.shape == (1, 1) and float(s.array) == 2.5


def test_dirac_and_functional_dispatch_host_semantics():
    """``DiracFunctional`` shapes (src/linpde_gp/linfunctls/_dirac.py:10-45: output = batch + codomain) and the argument
    checks of ``linfunctl(k, argnum)`` (covfuncs/linfunctls/_registry.py) that run before any device work."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.randprocs import covfuncs, crosscov

    X = np.linspace(0, 1, 12).reshape(3, 4)
    d = lg.linfunctls.DiracFunctional((), (), X)
    assert d.output_shape == (3, 4) and d.X_batch_shape == (3, 4) and d.X_batch_ndim == 2
    assert d.input_domain_shape == () and d.input_codomain_shape == ()
    np.testing.assert_array_equal(d(lg.functions.Constant((), 2.0)), np.full((3, 4), 2.0))
    with pytest.raises(ValueError):
        lg.linfunctls.DiracFunctional((2,), (), np.zeros((5, 3)))
    d2 = lg.linfunctls.DiracFunctional((2,), (), np.zeros((5, 2)))
    assert d2.output_shape == (5,)
    k = covfuncs.Matern((), nu=1.5, lengthscales=1.0)
    with pytest.raises(ValueError):
        d(k, argnum=2)
    z = d(covfuncs.Zero(()), argnum=1)  # the zero kernel needs no device work to be pushed through a functional
    assert isinstance(z, crosscov.Zero) and z.randvar_shape == (3, 4) and not z.reverse and z.randproc_input_shape == ()
    assert d(covfuncs.Zero(()), argnum=0).reverse
    assert isinstance(-z, crosscov.ScaledProcessVectorCrossCovariance) and (2.0 * z).scalar == 2.0
    assert isinstance(z + z, crosscov.SumProcessVectorCrossCovariance) and len((z + z + z).summands) == 3
    pv2 = lg.linfunctls.DiracFunctional((2,), (), np.zeros((5, 2)))(covfuncs.Zero((2,)), argnum=1)
    with pytest.raises(ValueError):
        pv2(np.zeros((4, 3)))  # trailing shape must equal the input shape of the process (checked before device work)
    with pytest.raises(ValueError):
        pv2.evaluate_linop(np.zeros((4, 3)))


def test_polynomial_functions_and_closed_form_operators():
    """``functions.Monomial / Polynomial / RationalPolynomial`` (src/linpde_gp/functions/_polynomial.py:17-238): Horner
    evaluation, exact calculus on the coefficients, arithmetic, ``//`` by a monomial; differential operators and
    ``LebesgueIntegral`` (functions/_linfunctls.py:9-13) apply to them in closed form."""
    from fractions import Fraction

    import linpde_gp_b200 as lg
    from linpde_gp_b200.functions import Constant, Monomial, Polynomial, RationalPolynomial
    from linpde_gp_b200.linfuncops import diffops

    x = np.linspace(-2, 2, 9).reshape(3, 3)
    p = Polynomial([1.0, -2.0, 0.0, 3.0])
    assert p.degree == 3 and p.coefficients == (1.0, -2.0, 0.0, 3.0) and p.input_shape == () and p.output_shape == ()
    np.testing.assert_allclose(p(x), 1 - 2 * x + 3 * x**3)
    np.testing.assert_allclose(Monomial(4)(x), x**4)
    assert p.differentiate().coefficients == (-2.0, 0.0, 9.0)
    assert p.integrate().coefficients == (0.0, 1.0, -1.0, 0.0, 0.75)
    assert (-p).coefficients == (-1.0, 2.0, -0.0, -3.0)
    assert (p + Polynomial([1.0, 1.0])).coefficients == (2.0, -1.0, 0.0, 3.0)
    assert (p - Polynomial([0.0, 0.0, 0.0, 3.0, 1.0])).coefficients == (1.0, -2.0, 0.0, 0.0, -1.0)
    assert (p + Constant((), 2.0)).coefficients == (3.0, -2.0, 0.0, 3.0)
    assert (0.5 * p).coefficients == (0.5, -1.0, 0.0, 1.5)
    assert (Polynomial([0.0, 0.0, 2.0, 5.0]) // Monomial(2)).coefficients == (2.0, 5.0)
    with pytest.raises(ValueError):
        p // Monomial(1)  # constant coefficient is not zero
    with pytest.raises(ValueError):
        p // Monomial(5)
    assert Polynomial([]).coefficients == (0.0,)
    with pytest.raises(ValueError):
        Monomial(-1)

    # the exact Matern-5/2 polynomial P_0 = 1 + r + r^2/3 and its derivative table entries
    q = RationalPolynomial([1, 1, Fraction(1, 3)])
    assert q.rational_coefficients == (Fraction(1), Fraction(1), Fraction(1, 3))
    assert q.differentiate().rational_coefficients == (Fraction(1), Fraction(2, 3))
    assert q.integrate().rational_coefficients == (Fraction(0), Fraction(1), Fraction(1, 2), Fraction(1, 9))
    assert isinstance(q - q.differentiate(), RationalPolynomial)
    assert (q - q.differentiate()).rational_coefficients == (Fraction(0), Fraction(1, 3), Fraction(1, 3))
    assert (3 * q).rational_coefficients == (Fraction(3), Fraction(3), Fraction(1))
    assert repr(q) == "1 + x^1 + 1/3 x^2" and repr(RationalPolynomial([0])) == "0"
    assert not isinstance(q + p, RationalPolynomial) and (q + p).coefficients[0] == 2.0
    np.testing.assert_allclose(q(x), 1 + x + x**2 / 3)

    # operators in closed form: L = -d^2/dx^2 + 2 d/dx on p
    L = -1.0 * diffops.Laplacian(()) + 2.0 * diffops.Derivative(1)
    Lp = L(p)
    np.testing.assert_allclose(Lp(x), -(18 * x) + 2 * (-2 + 9 * x**2))
    assert isinstance(diffops.Laplacian(())(Constant((), 3.0)), lg.functions.Zero)
    # prior-mean integrals: int_a^b p
    I = lg.linfunctls.LebesgueIntegral((-1.0, 2.0))
    P = p.integrate()
    assert abs(I(p) - (P(np.asarray(2.0)) - P(np.asarray(-1.0)))) < 1e-14
    assert abs(I(Constant((), 2.0)) - 6.0) < 1e-14


def test_domains_host_semantics():
    """``linpde_gp.domains`` (src/linpde_gp/domains/*.py): intervals, points, boxes, Cartesian products, boundaries,
    uniform grids (a box's grid is a TensorProductGrid, _box.py:81-114) and ``asdomain`` conversions."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200 import domains
    from linpde_gp_b200.randprocs.covfuncs import TensorProductGrid

    iv = domains.asdomain([-1.0, 1.0])
    assert isinstance(iv, domains.Interval) and iv.shape == () and tuple(iv) == (-1.0, 1.0) and len(iv) == 2
    assert iv[0] == -1.0 and iv[-1] == 1.0 and iv.volume == 2.0 and 0.3 in iv and 1.5 not in iv and np.zeros(2) not in iv
    assert iv == domains.Interval(-1, 1) and iv != domains.Interval(-1, 2)
    lo, hi = iv.boundary
    assert isinstance(lo, domains.Point) and float(lo) == -1.0 and float(hi) == 1.0 and lo.boundary == (lo,) and lo.volume == 0.0
    np.testing.assert_allclose(iv.uniform_grid(5, inset=0.1), np.linspace(-0.9, 0.9, 5))
    with pytest.raises(ValueError):
        domains.Interval(1.0, 0.0)
    with pytest.raises(ValueError):
        domains.asdomain([1.0, 2.0, 3.0])

    box = domains.asdomain([np.array([0.0, -1.0]), np.array([1.0, 1.0])])
    assert isinstance(box, domains.Box) and box.shape == (2,) and len(box) == 2 and box.volume == 2.0
    np.testing.assert_array_equal(box.bounds, [[0.0, 1.0], [-1.0, 1.0]])
    assert np.array([0.5, 0.0]) in box and np.array([1.5, 0.0]) not in box and box[1] == domains.Interval(-1, 1)
    assert box == domains.Box(np.array([[0.0, 1.0], [-1.0, 1.0]])) and isinstance(box[0:1], domains.Box)
    grid = box.uniform_grid((3, 5), inset=(0.0, 0.5))
    assert isinstance(grid, TensorProductGrid) and grid.shape == (3, 5, 2)
    np.testing.assert_allclose(grid.factors[0], [0.0, 0.5, 1.0])
    np.testing.assert_allclose(grid.factors[1], np.linspace(-0.5, 0.5, 5))
    assert box.uniform_grid(4).shape == (4, 4, 2)
    parts = box.boundary
    assert len(parts) == 4 and all(isinstance(p, domains.CartesianProduct) and p.shape == (2,) for p in parts)
    assert isinstance(parts[0][0], domains.Point) and float(parts[0][0]) == 0.0 and parts[0][1] == domains.Interval(-1, 1)
    edge = parts[3].uniform_grid(7)  # y = 1 edge: one point along the collapsed axis
    assert edge.shape == (7, 1, 2) and np.all(np.asarray(edge)[..., 1] == 1.0)
    with pytest.raises(ValueError):
        domains.Box(np.array([[1.0, 0.0]]))
    with pytest.raises(TypeError):
        domains.Box(np.array([[0, 1]]))
    cp = domains.CartesianProduct(domains.Interval(0.0, 5.0), iv)
    assert cp.shape == (2,) and cp.volume == 10.0 and cp.uniform_grid((4, 3)).shape == (4, 3, 2) and len(cp.boundary) == 4
    assert lg.domains.Point(np.array([1.0, 2.0])).shape == (2,)


def test_pde_problem_definitions():
    """``linpde_gp.problems.pde`` (src/linpde_gp/problems/pde/*.py): Poisson / heat Dirichlet problems, their operators,
    boundary conditions and analytic solutions."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200 import domains
    from linpde_gp_b200.linfuncops import diffops
    from linpde_gp_b200.problems import pde

    iv = domains.asdomain([-1.0, 1.0])
    bvp = pde.PoissonEquationDirichletProblem(iv, rhs=lg.functions.Constant((), 2.0), boundary_values=(0.0, 1.0), alpha=1.0)
    assert isinstance(bvp.pde.diffop, diffops.ScaledLinearDifferentialOperator) and bvp.domain == iv
    assert bvp.pde.diffop._terms() == {(2,): -1.0}
    X_bc, Y_bc = pde.get_1d_dirichlet_boundary_observations(bvp.boundary_conditions)
    np.testing.assert_array_equal(X_bc, [-1.0, 1.0])
    np.testing.assert_array_equal(Y_bc, [0.0, 1.0])
    xs = np.linspace(-1, 1, 11)
    u = bvp.solution(xs)
    h = 1e-4  # -u'' = 2 by central differences, boundary values reproduced
    np.testing.assert_allclose(-(bvp.solution(xs[1:-1] + h) - 2 * u[1:-1] + bvp.solution(xs[1:-1] - h)) / h**2, 2.0, rtol=1e-6)
    assert u[0] == 0.0 and abs(u[-1] - 1.0) < 1e-15
    sq = domains.asdomain([np.zeros(2), np.ones(2)])
    bvp2 = pde.PoissonEquationDirichletProblem(sq, rhs=lg.functions.Constant((2,), 2.0))
    assert len(bvp2.boundary_conditions) == 4 and bvp2.solution is None
    assert all(isinstance(bc.values, lg.functions.Zero) and bc.operator.input_domain_shape == (2,) for bc in bvp2.boundary_conditions)
    with pytest.raises(ValueError):
        pde.LinearPDE(iv, diffops.Laplacian((2,)))

    ibvp = pde.HeatEquationDirichletProblem(t0=0.0, T=5.0, spatial_domain=iv, alpha=0.1,
                                            initial_values=lg.functions.TruncatedSineSeries(iv, coefficients=[1.0, 2.0]))
    assert isinstance(ibvp.pde.diffop, diffops.HeatOperator) and ibvp.pde.diffop.alpha == 0.1
    assert ibvp.t0 == 0.0 and ibvp.T == 5.0 and ibvp.spatial_domain == iv and ibvp.temporal_domain == domains.Interval(0, 5)
    X_ic = ibvp.initial_domain.uniform_grid(5, inset=1e-6)
    assert X_ic.shape == (1, 5, 2) and np.all(np.asarray(X_ic)[..., 0] == 0.0)
    Y_ic = ibvp.initial_condition.values(np.asarray(X_ic)[..., 1])
    np.testing.assert_allclose(ibvp.solution(np.asarray(X_ic)), Y_ic, rtol=1e-13, atol=1e-15)  # u(t0, x) = initial values
    for bc in ibvp.boundary_conditions:
        Xb = bc.boundary.uniform_grid(50)
        assert Xb.shape == (50, 1, 2) and np.all(bc.values(np.asarray(Xb)) == 0.0)
        assert np.max(np.abs(ibvp.solution(np.asarray(Xb)))) < 1e-14  # zero Dirichlet values
    # the analytic solution solves the PDE: u_t - alpha u_xx = 0 by finite differences at interior points
    tx = np.asarray(ibvp.domain.uniform_grid((6, 7), inset=(0.5, 0.3)))
    e_t, e_x, h = np.array([1.0, 0.0]), np.array([0.0, 1.0]), 1e-4
    u_t = (ibvp.solution(tx + h * e_t) - ibvp.solution(tx - h * e_t)) / (2 * h)
    u_xx = (ibvp.solution(tx + h * e_x) - 2 * ibvp.solution(tx) + ibvp.solution(tx - h * e_x)) / h**2
    assert np.max(np.abs(u_t - 0.1 * u_xx)) < 1e-5
    assert isinstance(pde.HeatEquationDirichletProblem(0.0, iv, T=1.0).solution, lg.functions.Zero)


def test_api_descriptors_equal_spec_descriptors_for_every_reference_case():
    """The reference-style objects (kernel classes + operators through the dispatch) lower to the same device
    descriptor as the flat case specification -- all 117 frozen cases, radial family included (no GPU needed)."""
    import json
    import os

    from tests import helpers

    K = np.load(os.path.join(os.path.dirname(__file__), "golden", "kernels.npz"))
    specs = json.loads(bytes(K["__specs__"]).decode())
    assert len(specs) == 117
    for spec in specs:
        descs = covfuncs.device_descriptors(helpers.api_L0kL1(spec))
        ref = helpers.desc_from_spec(spec)
        assert len(descs) == 1, spec["name"]
        a, b = np.array(ref.coef[:]), np.array(descs[0].coef[:])
        assert np.allclose(a, b, rtol=1e-14, atol=0), spec["name"]
        assert abs(ref.diag_value - descs[0].diag_value) <= 1e-14 * max(1.0, abs(ref.diag_value)), spec["name"]


def test_linear_interpolation_basis_and_l2_projection_host_side():
    """``functions.bases.UnivariateLinearInterpolationBasis`` (src/linpde_gp/functions/bases/_fem.py:7-117) and the host
    parts of ``L2Projection_UnivariateLinearInterpolationBasis`` (linfunctls/projections/l2/_fem.py:14-62): sentinel
    nodes, partition of unity, support bounds, mass matrix, quadrature weights, argument errors."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.linfunctls.projections.l2 import L2Projection_UnivariateLinearInterpolationBasis

    grid = np.array([0.0, 0.2, 0.5, 0.6, 1.0])
    b = lg.functions.bases.UnivariateLinearInterpolationBasis(grid)
    assert len(b) == 5 and b.output_shape == (5,) and not b.zero_boundary
    assert np.allclose(b.grid, [-0.2, 0.0, 0.2, 0.5, 0.6, 1.0, 1.4])
    x = np.linspace(0.0, 1.0, 41)
    assert np.allclose(b(x).sum(-1), 1.0)                     # partition of unity on the grid
    assert np.allclose(b(grid), np.eye(5))                    # nodal basis
    assert b.support_bounds(0) == (0.0, 0.2) and b.support_bounds(4) == (0.6, 1.0) and b.support_bounds(2) == (0.2, 0.6)
    assert np.allclose(b.eval_elem(2, x), b(x)[:, 2])
    bz = lg.functions.bases.UnivariateLinearInterpolationBasis(grid, zero_boundary=True)
    assert len(bz) == 3 and np.allclose(bz(grid)[:, 0], [0, 1, 0, 0, 0])
    with pytest.raises(ValueError):
        lg.functions.bases.UnivariateLinearInterpolationBasis([0.0, 1.0])
    proj = b.l2_projection()
    assert isinstance(proj, L2Projection_UnivariateLinearInterpolationBasis) and proj.normalized
    assert proj.output_shape == (5,) and proj.input_shapes == ((), ())
    M = proj.mass_matrix()
    assert np.allclose(M, M.T) and np.isclose(M.sum(), 1.0)   # int (sum phi_i)(sum phi_j) = |domain|
    nodes, W = b.gauss_legendre(6)
    assert W.shape == (5, 24) and np.allclose(W.sum(1), M.sum(1)) and np.allclose(W @ nodes, M @ grid)
    assert [a[1] for a in proj._atoms()] == ["proj"]
    with pytest.raises(TypeError):
        L2Projection_UnivariateLinearInterpolationBasis(grid)
