"""Block linear operators on the device (seam 2 of 8b): the reference's own tests for
``BlockMatrix`` / ``BlockMatrix2x2`` / ``ConcatenatedLinearOperator`` (tests/linpde_gp/linops/test_block.py,
test_symmetric_block.py) restated against this package, plus the parts of ``BlockMatrix2x2`` the conditioning code
calls (``schur``, ``L_A_inv_B``, ``schur_update``, block-triangular solves, ``det``; _block.py:191-292).

Tolerances: every comparison is against numpy/scipy FP64 on the same inputs at 1e-10 relative (the matrices are
small and well conditioned except where noted)."""
import numpy as np
import pytest
import scipy.linalg

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _expquad(x):
    return np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2)


def _spd(M):
    from linpde_gp_b200 import linops

    op = linops.Matrix(M)
    op.is_symmetric = True
    op.is_positive_definite = True
    return op


def _sbm(K, cut):
    from linpde_gp_b200 import linops

    return linops.BlockMatrix2x2(_spd(K[:cut, :cut]), linops.Matrix(K[:cut, cut:]), None, _spd(K[cut:, cut:]), is_spd=True)


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.max(np.abs(a - b)) <= rtol * max(np.max(np.abs(b)), 1e-300), np.max(np.abs(a - b))


# -- test_symmetric_block.py ------------------------------------------------------------------------------------------
def test_cholesky_and_inverse_3x3():
    """test_symmetric_block.py:41-53 (ExpQuad on x = 1, 2, 3, blocks 2 + 1)."""
    K = _expquad(np.array([1.0, 2.0, 3.0]))
    sbm, full = _sbm(K, 2), _spd(K)
    L_full, L_block = full.cholesky(), sbm.cholesky()
    assert L_full.shape == (3, 3) and L_block.shape == (3, 3)
    v = np.array([10.0, 11.0, 12.0])
    _close(L_block @ v, np.linalg.cholesky(K) @ v)
    _close(L_full @ v, L_block @ v)
    _close(sbm.inv() @ v, np.linalg.solve(K, v))
    _close(full.inv() @ v, sbm.inv() @ v)


def test_cholesky_and_inverse_nested_5x5():
    """test_symmetric_block.py:56-107 (nested SBM: (2 + 2) + 1)."""
    from linpde_gp_b200 import linops

    K = _expquad(np.arange(1.0, 6.0))
    inner = _sbm(K[:4, :4], 2)
    F = _spd(K[4:, 4:])
    sbm = linops.BlockMatrix2x2(inner, linops.Matrix(K[:4, 4:]), None, F, is_spd=True)
    v = np.array([10.0, 11.0, 12.0, 13.0, 14.0])
    L = sbm.cholesky()
    assert L.shape == (5, 5)
    _close(L.todense(), np.linalg.cholesky(K))
    _close(L @ v, np.linalg.cholesky(K) @ v)
    _close(sbm.inv() @ v, np.linalg.solve(K, v))
    _close(sbm.todense(), K)
    assert inner.cholesky() is inner.cholesky()  # the inner factor is cached and was reused, not refactored


@pytest.mark.parametrize("n,cut", [(37, 20), (300, 129), (513, 256), (260, 1)])
def test_spd_block_quantities(n, cut):
    """schur / L_A_inv_B / schur_update / solve / det of an SPD block matrix (_block.py:191-292) against numpy, at sizes
    that cross leaf (128) and odd-padding boundaries."""
    rng = np.random.default_rng(n)
    x = np.sort(rng.uniform(0, 0.05 * n, n))
    K = _expquad(x) + 1e-3 * np.eye(n)
    sbm = _sbm(K, cut)
    A, B, D = K[:cut, :cut], K[:cut, cut:], K[cut:, cut:]
    LA = np.linalg.cholesky(A)
    LAinvB = scipy.linalg.solve_triangular(LA, B, lower=True)
    S = D - LAinvB.T @ LAinvB
    _close(sbm.L_A_inv_B.todense(), LAinvB, 1e-9)
    _close(sbm.schur.todense(), S, 1e-9)
    assert sbm.schur.is_symmetric and sbm.schur.is_positive_definite
    L = sbm.cholesky(True)
    Ld = L.todense()
    assert np.allclose(Ld, np.tril(Ld)) and np.all(np.diag(Ld) > 0)
    _close(Ld @ Ld.T, K, 1e-11)
    _close(sbm.cholesky(False).todense(), Ld.T)
    u, v = rng.standard_normal(cut), rng.standard_normal(n - cut)
    sol = np.linalg.solve(K, np.concatenate([u, v]))
    scale = np.max(np.abs(sol))
    got = sbm.schur_update(np.linalg.solve(A, u), v)
    assert np.max(np.abs(got - sol)) <= 1e-8 * scale
    Bm = rng.standard_normal((n, 3))
    X = sbm.solve(Bm)
    assert np.max(np.abs(K @ X - Bm)) <= 1e-8 * np.max(np.abs(Bm)) * np.linalg.cond(K) ** 0.5
    x1 = sbm.solve(Bm[:, 0])
    assert np.max(np.abs(x1 - X[:, 0])) <= 1e-12 * np.max(np.abs(X))
    assert abs(sbm.logabsdet() - np.linalg.slogdet(K)[1]) <= 1e-9 * abs(np.linalg.slogdet(K)[1]) + 1e-9
    assert abs(sbm.trace() - np.trace(K)) <= 1e-12 * np.trace(K)
    # triangular solves with the bordered factor and its transpose (scipy solve_triangular, pn _linear_operator.py:296-299)
    Lref = np.linalg.cholesky(K)
    y = L.inv() @ Bm
    assert np.max(np.abs(Lref @ y - Bm)) <= 1e-9 * np.max(np.abs(Bm)) * np.linalg.cond(Lref)
    z = L.T.inv() @ Bm[:, 1]
    assert np.max(np.abs(Lref.T @ z - Bm[:, 1])) <= 1e-9 * np.max(np.abs(Bm)) * np.linalg.cond(Lref)
    assert abs(L.det() - np.exp(0.5 * np.linalg.slogdet(K)[1])) <= 1e-9 * L.det()


def test_block_diagonal_and_triangular_structures():
    """Block-diagonal (B = C = None) and block-triangular BlockMatrix2x2 (_block.py:131-175, 244-266)."""
    from linpde_gp_b200 import linops

    rng = np.random.default_rng(5)
    K = _expquad(np.linspace(0, 3, 9)) + 1e-2 * np.eye(9)
    A, D = K[:5, :5], K[5:, 5:]
    bd = linops.BlockMatrix2x2(_spd(A), None, None, _spd(D))
    assert bd.is_block_diagonal and bd.is_symmetric
    dense = scipy.linalg.block_diag(A, D)
    _close(bd.todense(), dense)
    b = rng.standard_normal(9)
    _close(bd.solve(b), np.linalg.solve(dense, b), 1e-9)
    _close(bd.schur_update(np.linalg.solve(A, b[:5]), b[5:]), np.linalg.solve(dense, b), 1e-9)
    assert abs(bd.det() - np.linalg.det(dense)) <= 1e-9 * abs(np.linalg.det(dense))
    # lower block-triangular from two Cholesky factors and a dense coupling block
    LA, LD = _spd(A).cholesky(True), _spd(D).cholesky(True)
    C = rng.standard_normal((4, 5))
    lt = linops.BlockMatrix2x2(LA, None, linops.Matrix(C), LD)
    assert lt.is_lower_triangular and not lt.is_symmetric
    dense_lt = np.block([[np.linalg.cholesky(A), np.zeros((5, 4))], [C, np.linalg.cholesky(D)]])
    _close(lt.todense(), dense_lt)
    Bm = rng.standard_normal((9, 2))
    _close(lt.solve(Bm), scipy.linalg.solve_triangular(dense_lt, Bm, lower=True), 1e-9)
    _close(lt.inv() @ b, scipy.linalg.solve_triangular(dense_lt, b, lower=True), 1e-9)
    ut = lt.T
    assert ut.is_upper_triangular
    _close(ut.todense(), dense_lt.T)
    _close(ut.solve(b), scipy.linalg.solve_triangular(dense_lt.T, b, lower=False), 1e-9)
    assert abs(lt.det() - np.linalg.det(dense_lt)) <= 1e-9 * abs(np.linalg.det(dense_lt))
    # general 2x2 block matrix: products / transposes only
    g = linops.BlockMatrix2x2(linops.Matrix(A), linops.Matrix(C.T), linops.Matrix(C), linops.Matrix(D))
    dense_g = np.block([[A, C.T], [C, D]])
    _close(g @ b, dense_g @ b)
    _close(g.T.todense(), dense_g.T)
    with pytest.raises(ValueError):
        linops.BlockMatrix2x2(_spd(A), linops.Matrix(C.T), linops.Matrix(C), _spd(D), is_spd=True)
    with pytest.raises(ValueError):
        linops.BlockMatrix2x2(linops.Matrix(A), linops.Matrix(C), None, linops.Matrix(D))


def test_not_positive_definite_block_raises():
    from linpde_gp_b200 import linops

    K = _expquad(np.array([0.0, 1.0, 2.0, 3.0]))
    K[2, 3] = K[3, 2] = 1.5  # |correlation| > 1: the Schur complement of the last row is negative
    sbm = linops.BlockMatrix2x2(_spd(K[:2, :2]), linops.Matrix(K[:2, 2:]), None, _spd(K[2:, 2:]), is_spd=True)
    with pytest.raises(np.linalg.LinAlgError):
        sbm.cholesky()
    assert sbm.is_positive_definite is False


# -- test_block.py ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [2349023, 896934, 12983, 5492538])
def test_block_matrix_matmul_todense_transpose(seed):
    """test_block.py:9-56: random grids of Matrix / Zero / Identity blocks."""
    from linpde_gp_b200 import linops

    rng = np.random.RandomState(seed)
    shape = (rng.randint(5) + 1, rng.randint(5) + 1)
    row_dims = [rng.randint(3) + 1 for _ in range(shape[0])]
    col_dims = [rng.randint(3) + 1 for _ in range(shape[1])]

    def pull(i, j):
        shp = (row_dims[i], col_dims[j])
        choices = [linops.Matrix(rng.rand(*shp)), linops.Zero(shp)]
        if shp[0] == shp[1]:
            choices.append(linops.Identity(shp))
        return choices[rng.randint(len(choices))]

    blocks = [[pull(i, j) for j in range(shape[1])] for i in range(shape[0])]
    dense = np.block([[b.todense() for b in row] for row in blocks])
    op = linops.BlockMatrix(blocks)
    assert op.shape == dense.shape
    x = rng.rand(op.shape[1])
    np.testing.assert_allclose(op @ x, dense @ x, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(op.todense(), dense)
    np.testing.assert_allclose(op.T.todense(), dense.T)
    with pytest.raises(ValueError):
        linops.BlockMatrix([[linops.Zero((2, 2)), linops.Zero((3, 2))]])


@pytest.mark.parametrize("axis", [0, 1, -1, -2])
def test_concatenated_linear_operator(axis):
    """src/linpde_gp/linops/_concatenated.py:8-72."""
    from linpde_gp_b200 import linops

    rng = np.random.default_rng(3)
    ax = axis % 2
    mats = [rng.standard_normal((4, 3) if ax == 1 else (3, 4)), rng.standard_normal((4, 5) if ax == 1 else (5, 4)),
            rng.standard_normal((4, 1) if ax == 1 else (1, 4))]
    op = linops.ConcatenatedLinearOperator(tuple(mats), axis=axis)
    dense = np.concatenate(mats, axis=ax)
    assert op.shape == dense.shape and op.axis == ax and len(op.linops) == 3
    np.testing.assert_allclose(op.todense(), dense)
    x = rng.standard_normal((dense.shape[1], 2))
    np.testing.assert_allclose(op @ x, dense @ x, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(op.T.todense(), dense.T)
    with pytest.raises(ValueError):
        linops.ConcatenatedLinearOperator((), axis=0)
    with pytest.raises(ValueError):
        linops.ConcatenatedLinearOperator(tuple(mats), axis=2)


def test_covariance_blocks_reuse_cached_factor():
    """The conditioning pattern of _conditional.py:253-294 spelled out with public linops: Gram of batch 1 factored once,
    batch 2 bordered onto it with covariance linops as blocks (assembled by the Gram kernel straight into the factor)."""
    from linpde_gp_b200 import linops
    from linpde_gp_b200.randprocs import covfuncs

    k = 2.0 * covfuncs.Matern((), nu=2.5, lengthscales=0.4)
    X1, X2 = np.linspace(-1, 1, 131), np.linspace(-0.93, 0.97, 57)
    A = k.linop(X1)
    A.is_positive_definite = True
    LA = A.cholesky(True)
    D = k.linop(X2)
    sbm = linops.BlockMatrix2x2(A, k.linop(X1, X2), None, D, is_spd=True)
    K = k.matrix(np.concatenate([X1, X2]))
    L = sbm.cholesky(True)
    assert A.cholesky(True) is LA
    Ld = L.todense()
    assert np.max(np.abs(Ld @ Ld.T - K)) <= 1e-12 * np.max(np.abs(K))
    b = np.random.default_rng(0).standard_normal(len(K))
    x = sbm.solve(b)
    assert np.max(np.abs(K @ x - b)) <= 1e-8 * np.max(np.abs(x)) * np.max(np.abs(K))


def test_kronecker_structured_solve_cholesky_det():
    """``Kronecker.solve / inv / cholesky / logabsdet / det / trace`` (pn/linops/_kronecker.py:122-166, 233-242) never form
    the (n1 n2)^2 matrix: against numpy on the dense Kronecker product, odd factor sizes (padded device factors),
    vectors, matrices and stacks of right-hand sides."""
    from linpde_gp_b200 import linops

    rng = np.random.default_rng(3)
    n1, n2 = 37, 50
    A0 = rng.standard_normal((n1, n1 + 5))
    B0 = rng.standard_normal((n2, n2 + 5))
    A, B = A0 @ A0.T / n1 + 0.3 * np.eye(n1), B0 @ B0.T / n2 + 0.3 * np.eye(n2)
    opA, opB = linops.Matrix(A), linops.Matrix(B)
    opA.is_symmetric = opB.is_symmetric = True
    K = linops.Kronecker(opA, opB)
    dense = np.kron(A, B)
    assert K.is_symmetric
    for rhs in (rng.standard_normal(n1 * n2), rng.standard_normal((n1 * n2, 3)), rng.standard_normal((2, n1 * n2, 5))):
        x = K.solve(rhs)
        ref = np.linalg.solve(dense, rhs)
        assert x.shape == ref.shape
        assert np.max(np.abs(x - ref)) <= 1e-10 * np.max(np.abs(ref))
    x = K.inv() @ rng.standard_normal(n1 * n2)
    assert x.shape == (n1 * n2,)
    L = K.cholesky()
    assert isinstance(L, linops.Kronecker) and L.is_lower_triangular
    Lref = np.linalg.cholesky(dense)
    assert np.max(np.abs(L.todense() - Lref)) <= 1e-11 * np.max(np.abs(Lref))
    b = rng.standard_normal((n1 * n2, 6))
    y = L.solve(b)  # triangular Kronecker: L_A^{-1} (x) L_B^{-1}
    assert np.max(np.abs(y - np.linalg.solve(Lref, b))) <= 1e-9 * np.max(np.abs(y))
    yt = L.T.solve(b)
    assert np.max(np.abs(yt - np.linalg.solve(Lref.T, b))) <= 1e-9 * np.max(np.abs(yt))
    sign, logdet = np.linalg.slogdet(dense)
    assert sign > 0 and abs(K.logabsdet() - logdet) <= 1e-9 * abs(logdet)
    assert abs(K.trace() - np.trace(dense)) <= 1e-10 * np.trace(dense)
    Ks = linops.Kronecker(linops.Matrix(A[:5, :5] * 1.0), linops.Matrix(B[:4, :4] * 1.0))
    Ks.A.is_symmetric = Ks.B.is_symmetric = True
    assert abs(Ks.det() - np.linalg.det(np.kron(A[:5, :5], B[:4, :4]))) <= 1e-9 * abs(Ks.det())


def test_kronecker_gram_of_a_gridded_product_kernel_solves_beyond_dense_size():
    """A product kernel on a 700 x 900 tensor grid (N = 630,000: the dense Gram would be 3.2 TB): ``k.linop(grid)`` stays
    a Kronecker product and ``solve`` runs on the two small factors; checked through the residual applied by ``@``."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200 import linops
    from linpde_gp_b200.randprocs import covfuncs

    g1, g2 = np.linspace(0.0, 1.0, 700), np.linspace(0.0, 2.0, 900)
    grid = covfuncs.TensorProductGrid(g1, g2)
    k = covfuncs.TensorProduct(covfuncs.Matern((), nu=1.5, lengthscales=0.05), covfuncs.Matern((), nu=0.5, lengthscales=0.1))
    K = k.linop(grid)
    assert isinstance(K, linops.Kronecker) and K.shape == (630000, 630000)
    rng = np.random.default_rng(0)
    b = rng.standard_normal(630000)
    x = K.solve(b)
    r = K @ x - b
    # backward error: the factors are fine-grid Matern Gram matrices (condition number ~1e7), so |x| >> |b|
    normK = np.abs(K.A.todense()).sum(1).max() * np.abs(K.B.todense()).sum(1).max()
    assert np.max(np.abs(r)) <= 1e-13 * normK * np.max(np.abs(x)), (np.max(np.abs(r)), normK, np.max(np.abs(x)))


@pytest.mark.parametrize("with_transform,with_noise", [(True, True), (True, False), (False, True)])
def test_condition_normal_on_observations_matches_numpy(with_transform, with_noise):
    """``randvars.condition_normal_on_observations`` (src/linpde_gp/randvars/_normal.py:8-71) against the same formulas
    in numpy / scipy (``cho_solve``), 1e-10 of the largest entry; also as ``Normal.condition_on_observations`` and with a
    single-row transform."""
    from linpde_gp_b200 import randvars

    rng = np.random.default_rng(5 + 2 * with_transform + with_noise)
    n, m = 40, 12
    R = rng.standard_normal((n, n))
    Sigma, mu = R @ R.T / n + 0.1 * np.eye(n), rng.standard_normal(n)
    A = rng.standard_normal((m, n)) if with_transform else None
    k = m if with_transform else n
    noise = randvars.Normal(rng.standard_normal(k), np.diag(rng.uniform(0.05, 0.2, k))) if with_noise else None
    y = rng.standard_normal(k)

    def numpy_posterior(A_, y_, noise_):
        cross = Sigma if A_ is None else A_ @ Sigma
        pm = mu if A_ is None else A_ @ mu
        pc = Sigma if A_ is None else cross @ A_.T
        if noise_ is not None:
            pm, pc = pm + noise_.mean.reshape(-1), pc + noise_.dense_cov
        gain = scipy.linalg.cho_solve((scipy.linalg.cholesky(pc, lower=True), True), cross).T
        return mu + gain @ (y_ - pm), Sigma - cross.T @ gain.T

    prior = randvars.Normal(mu, Sigma)
    post = randvars.condition_normal_on_observations(prior, y, noise, A)
    m_ref, C_ref = numpy_posterior(A, y, noise)
    assert np.max(np.abs(post.mean - m_ref)) <= 1e-10 * np.max(np.abs(m_ref))
    assert np.max(np.abs(post.dense_cov - C_ref)) <= 1e-10 * np.max(np.abs(Sigma))
    post2 = prior.condition_on_observations(y, noise, A)
    assert np.array_equal(post2.mean, post.mean)
    if with_transform:  # one scalar observation through a single row (the reference's 1-D `transform`)
        eps = randvars.Normal(np.asarray(0.3), np.asarray(0.01).reshape(1, 1)) if with_noise else None
        post1 = randvars.condition_normal_on_observations(prior, np.asarray(0.7), eps, A[0])
        eps1 = randvars.Normal(np.asarray([0.3]), np.asarray([[0.01]])) if with_noise else None
        m1, C1 = numpy_posterior(A[:1], np.asarray([0.7]), eps1)
        assert np.max(np.abs(post1.mean - m1)) <= 1e-10 * np.max(np.abs(m1))
        assert np.max(np.abs(post1.dense_cov - C1)) <= 1e-10 * np.max(np.abs(Sigma))
