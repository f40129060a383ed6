"""GPU parity tests through the reference-style Python API (-> ctypes -> liblpgp.so): kernels against the frozen
outputs of the real reference, GP conditioning / posterior against the reference goldens and the oracle."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from tests import helpers
from tests.golden import cases as gcases

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
_K = np.load(os.path.join(GOLDEN, "kernels.npz"))
SPECS = json.loads(bytes(_K["__specs__"]).decode())
GRAM_TOL = 1e-12   # north_star: Gram entries rel err <= 1e-12
POST_TOL = 1e-8    # north_star: posterior mean/cov rel err <= 1e-8 after factorisation


# every reference case lowers to a device descriptor (product form, or the radial family for the isotropic
# multi-dimensional Matern kernels)
LOWERABLE = SPECS


@pytest.mark.parametrize("spec", LOWERABLE, ids=[s["name"] for s in LOWERABLE])
def test_L0kL1_matrix_and_call(spec):
    """Analogue of test_diffops.py::test_L0kL1 with the frozen reference output as the oracle."""
    k = helpers.api_L0kL1(spec)
    shape = gcases.kernel_input_shape(spec["kernel"])
    X = gcases.sobol_points(shape)
    K_ref, d_ref = _K[spec["name"] + "__K"], _K[spec["name"] + "__diag"]
    scale = np.max(np.abs(K_ref))
    K = k.matrix(X[:32], X)
    assert K.shape == (32, 128) and K.dtype == np.float64
    assert np.max(np.abs(K - K_ref)) <= GRAM_TOL * scale
    Kc = k(X[:32, None], X[None, :])          # numpy-broadcast call
    assert np.max(np.abs(Kc - K_ref)) <= GRAM_TOL * scale
    assert np.max(np.abs(k(X[:32], None) - d_ref)) <= GRAM_TOL * scale
    pairs = k(X[:32], X[32:64])               # element-wise pairs
    full = k.matrix(X[:32], X[32:64])
    assert np.max(np.abs(pairs - np.diag(full))) <= GRAM_TOL * scale


def test_symmetric_matrix_and_linop_matmul():
    spec = next(s for s in SPECS if s["name"] == "ns_poisson2d_LkL")
    k = helpers.api_L0kL1(spec)
    X = gcases.sobol_points((2,))
    G = k.matrix(X)
    assert np.allclose(G, G.T, rtol=0, atol=1e-13 * np.max(np.abs(G)))
    assert np.max(np.abs(G[:32] - _K[spec["name"] + "__K"])) <= GRAM_TOL * np.max(np.abs(G))
    op = k.linop(X)
    v = np.random.default_rng(0).standard_normal((128, 3))
    assert np.max(np.abs(op @ v - G @ v)) <= 1e-11 * np.max(np.abs(G @ v))
    assert np.max(np.abs(op @ v[:, 0] - G @ v[:, 0])) <= 1e-11 * np.max(np.abs(G @ v))


def test_batch_shapes_are_flattened_c_order():
    k = helpers.api_kernel({"scale": 2.0, "base": {"kind": "expquad", "input_shape": [2], "lengthscales": [0.7, 1.3]}})
    X = np.random.default_rng(1).uniform(size=(4, 5, 2))
    assert np.allclose(k.matrix(X), k.matrix(X.reshape(-1, 2)), rtol=0, atol=1e-15)
    assert k(X, None).shape == (4, 5)


GP_FILES = sorted(glob.glob(os.path.join(GOLDEN, "gp_*.npz")))


@pytest.mark.parametrize("path", GP_FILES, ids=[os.path.basename(p)[3:-4] for p in GP_FILES])
def test_conditioning_matches_reference_golden(path):
    g = np.load(path)
    problem = json.loads(bytes(g["problem"]).decode())
    post, res = helpers.api_solve(problem)
    gs = np.max(np.abs(g["gram"]))
    assert np.max(np.abs(res["gram"] - g["gram"])) <= 1e-11 * gs   # G reconstructed as L L^T
    for key in ("mean", "var", "cov"):
        sc = max(np.max(np.abs(g[key])), np.max(np.abs(g["var"])))
        assert np.max(np.abs(res[key] - g[key])) <= POST_TOL * sc, key
    # representer weights: looser (cond(G) amplifies), but the fit they produce is pinned above
    assert np.max(np.abs(res["w"] - g["w"])) <= 1e-6 * np.max(np.abs(g["w"]))
    assert len(post._Ys) == len(problem["blocks"])


def test_iterative_equals_batch_conditioning():
    """tests/linpde_gp/randprocs/test_posterior_gp.py:152-162."""
    import linpde_gp_b200 as lg
    from oracle import gp as ogp

    prob = ogp.golden_problems()["expquad_iterative"]
    post, it = helpers.api_solve(prob)
    X = np.concatenate([np.asarray(b["X"]) for b in prob["blocks"]])
    Y = np.concatenate([np.asarray(b["Y"]) for b in prob["blocks"]])
    nv = np.concatenate([np.full(len(b["Y"]), b["noise_var"] or 0.0) for b in prob["blocks"]])
    prior = lg.GaussianProcess(lg.functions.Zero(()), helpers.api_kernel(prob["kernel"]))
    one = prior.condition_on_observations(Y, X=X, b=lg.randvars.Normal(np.zeros(11), lg.linops.Scaling(nv)))
    Xt = np.asarray(prob["Xt"])
    np.testing.assert_allclose(one.mean(Xt), it["mean"], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(one.var(Xt), it["var"], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(one.cov.matrix(Xt[:16]), it["cov"], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(one.std(Xt) ** 2, it["var"], rtol=1e-7, atol=1e-10)
    marg = one(Xt[:16])
    np.testing.assert_allclose(marg.mean, it["mean"][:16], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(marg.dense_cov, it["cov"], rtol=1e-7, atol=1e-10)


def test_medium_poisson2d_against_oracle():
    """C2-shaped problem scaled to N = 2,304 (five batches): posterior within 1e-8 of the CPU oracle."""
    from oracle import gp as ogp

    prob = ogp.poisson2d_problem(2048, 64, seed=11, grid=24)
    ref = ogp.solve(prob)
    post, res = helpers.api_solve(prob)
    for key in ("mean", "var", "cov"):
        sc = max(np.max(np.abs(ref[key])), np.max(np.abs(ref["var"])))
        assert np.max(np.abs(res[key] - ref[key])) <= POST_TOL * sc, key
    # boundary values are reproduced and the PDE residual of the posterior mean vanishes at the collocation points
    Xb = np.asarray(prob["blocks"][0]["X"])
    assert np.max(np.abs(post.mean(Xb))) <= 1e-6
    from linpde_gp_b200.linfuncops import diffops

    Lpost = (-1.0 * diffops.Laplacian((2,)))(post)
    Xp = np.asarray(prob["blocks"][-1]["X"])[:200]
    assert np.max(np.abs(Lpost.mean(Xp) - 2.0)) <= 1e-6
    assert np.max(np.abs(Lpost.var(Xp))) <= 1e-6 * np.max(np.abs(Lpost._prior.cov(Xp[:1], None)))


def test_heat_medium_against_oracle():
    from oracle import gp as ogp

    prob = ogp.heat_problem(n_ic=33, n_bc=50, nt=40, nx=24, alpha=0.1, grid=20)
    ref = ogp.solve(prob)
    _, res = helpers.api_solve(prob)
    for key in ("mean", "var", "cov"):
        sc = max(np.max(np.abs(ref[key])), np.max(np.abs(ref["var"])))
        assert np.max(np.abs(res[key] - ref[key])) <= POST_TOL * sc, key


def test_not_positive_definite_raises_linalgerror():
    import linpde_gp_b200 as lg
    from linpde_gp_b200.randprocs import covfuncs

    prior = lg.GaussianProcess(lg.functions.Zero(()), covfuncs.ExpQuad((), lengthscales=1.0))
    X = np.array([0.0, 0.0, 1.0, 2.0])  # duplicated point -> singular Gram, no nugget is ever added
    with pytest.raises(np.linalg.LinAlgError):
        prior.condition_on_observations(np.zeros(4), X=X)
    op = covfuncs.ExpQuad(()).linop(X)
    with pytest.raises(np.linalg.LinAlgError):
        op.cholesky()
    assert op.is_positive_definite is False


def test_linop_solve_and_cholesky():
    """probnum tests/test_linops/test_linop_decompositions.py:17-62 on a covariance linop."""
    from linpde_gp_b200.randprocs import covfuncs

    k = 4.0 * covfuncs.Matern((), nu=2.5, lengthscales=0.5)
    X = np.linspace(-1, 1, 37)
    op = k.linop(X)
    G = op.todense()
    L = op.cholesky(lower=True)
    Ld = L.todense()
    assert np.allclose(Ld @ Ld.T, G, rtol=0, atol=1e-12 * np.max(np.abs(G)))
    assert np.all(np.diag(Ld) > 0) and np.allclose(Ld, np.tril(Ld))
    assert op.cholesky(True) is L  # cached
    U = op.cholesky(lower=False).todense()
    assert np.allclose(U, Ld.T)
    b = np.random.default_rng(0).standard_normal((37, 4))
    x = op.solve(b)
    assert np.max(np.abs(G @ x - b)) <= 1e-8 * np.max(np.abs(b))
    x1 = op.solve(b[:, 0])
    assert np.allclose(x1, x[:, 0], rtol=1e-9, atol=1e-12)


def test_one_shot_batches_equal_sequential_conditioning():
    """from_observation_batches (whole Gram assembled + factored at once) == batch-by-batch conditioning."""
    import linpde_gp_b200 as lg
    from oracle import gp as ogp

    for prob in (ogp.poisson2d_problem(700, 33, seed=4, grid=12, noise_bc=1e-6), ogp.heat_problem(n_ic=9, n_bc=21, nt=20, nx=11)):
        post_seq, seq = helpers.api_solve(prob)
        shape = gcases.kernel_input_shape(prob["kernel"])
        prior = lg.GaussianProcess(lg.functions.Zero(input_shape=shape), helpers.api_kernel(prob["kernel"]))
        batches = []
        for blk in prob["blocks"]:
            Y = np.asarray(blk["Y"], dtype=float)
            b = None
            if blk.get("noise_var") is not None:
                b = lg.randvars.Normal(np.zeros_like(Y), lg.linops.Scaling(np.broadcast_to(blk["noise_var"], Y.shape).copy()))
            batches.append((Y, np.asarray(blk["X"], dtype=float), helpers.api_op(blk["L"]), b))
        post = lg.ConditionalGaussianProcess.from_observation_batches(prior, batches)
        Xt = np.asarray(prob["Xt"], dtype=float)
        sc = np.max(np.abs(seq["var"]))
        assert np.max(np.abs(post.mean(Xt) - seq["mean"])) <= 1e-9 * max(sc, np.max(np.abs(seq["mean"])))
        assert np.max(np.abs(post.var(Xt) - seq["var"])) <= 1e-9 * sc
        assert np.max(np.abs(post.gram.todense() - seq["gram"])) <= 1e-11 * np.max(np.abs(seq["gram"]))
        # and the one-shot posterior can still be extended by bordering
        Xn = Xt[5:7] + 0.01234  # two new points that do not coincide with existing observations
        post2 = post.condition_on_observations(np.array([0.1, 0.2]), X=Xn)
        assert np.max(np.abs(post2.mean(Xn) - np.array([0.1, 0.2]))) <= 1e-6


def test_distributed_cholesky_single_rank_device_ops():
    """The multi-GPU code path with world_size 1 (no collectives) on the real CUDA kernels, ragged last block."""
    import torch

    from linpde_gp_b200.distributed import DistributedCholesky

    n = 1664 + 130
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(n, n + 8, dtype=torch.float64, device="cuda", generator=g)
    G = X @ X.T / n + 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
    ch = DistributedCholesky(n, nb=512)
    for i in ch.layout.local_blocks(0):
        lo, hi = ch.layout.block_bounds(i)
        ch.local_block_rows(i)[:, :hi].copy_(G[lo:hi, :hi])
    from linpde_gp_b200 import backend

    L_full = backend.alloc_matrix(n, n)
    L_full.zero_()
    ch.factor(L_full)
    L_ref = torch.linalg.cholesky(G)
    assert (torch.tril(ch.A_loc) - L_ref).abs().max().item() <= 1e-11 * L_ref.abs().max().item()
    assert (torch.tril(L_full) - L_ref).abs().max().item() <= 1e-11 * L_ref.abs().max().item()
    # inverted leaves of the whole factor are available for the triangular solves that follow
    W0 = ch.dinv[: 128 * 128].view(128, 128)
    assert (W0 @ L_ref[:128, :128] - torch.eye(128, dtype=torch.float64, device="cuda")).abs().max().item() <= 1e-10
    # not positive definite: LinAlgError with the position of the failing leading minor, no host sync inside the loop
    G2 = G.clone()
    G2[700, 700] = -1.0
    ch2 = DistributedCholesky(n, nb=512)
    for i in ch2.layout.local_blocks(0):
        lo, hi = ch2.layout.block_bounds(i)
        ch2.local_block_rows(i)[:, :hi].copy_(G2[lo:hi, :hi])
    with pytest.raises(np.linalg.LinAlgError, match="701-th"):
        ch2.factor()


def test_one_shot_with_distributed_factor_not_replicated():
    """replicate=False keeps the factor in its block-row layout (the N = 131,072 code path): representer weights by
    the owner-computes substitution, variance / covariance by streaming block rows of L -- here with world_size 1
    (no collectives) against sequential conditioning with the local factor."""
    import linpde_gp_b200 as lg
    from oracle import gp as ogp

    prob = ogp.poisson2d_problem(1500, 40, seed=9, grid=20)
    post_seq, seq = helpers.api_solve(prob)
    prior = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), helpers.api_kernel(prob["kernel"]))
    batches = [(np.asarray(b["Y"], dtype=float), np.asarray(b["X"], dtype=float), helpers.api_op(b["L"])) for b in prob["blocks"]]
    post = lg.ConditionalGaussianProcess.from_observation_batches(prior, batches, nb=256, replicate=False)
    Xt = np.asarray(prob["Xt"], dtype=float)
    sc = max(np.max(np.abs(seq["var"])), np.max(np.abs(seq["mean"])))
    assert np.max(np.abs(post.mean(Xt) - seq["mean"])) <= 1e-9 * sc
    assert np.max(np.abs(post.var(Xt) - seq["var"])) <= 1e-9 * sc
    assert np.max(np.abs(post.cov.matrix(Xt[:50]) - post_seq.cov.matrix(Xt[:50]))) <= 1e-9 * sc
    with pytest.raises(NotImplementedError):
        post.condition_on_observations(np.zeros(2), X=Xt[:2] + 0.0123)


def test_inverse_rhs_conditioning_lazy_kernel_noise():
    """Uncertain right-hand side (experiments/0003_poisson_1d_inverse_rhs.ipynb cell 19): ``b = -f_prior(X)`` makes
    the Gram matrix  L k_u L*(X, X) + k_f(X, X)  (_conditional.py:392-394).  The marginal's covariance stays a lazy
    operator and is accumulated into the Gram rows on the device; result == dense-``b`` path == numpy formula built
    from the oracle's matrices, for sequential, one-shot and distributed-layout conditioning."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.linfuncops import diffops
    from linpde_gp_b200.randprocs import covfuncs
    from oracle import covfuncs as ocf

    rng = np.random.default_rng(3)
    ell = 0.25
    ku_spec = {"scale": 4.0, "base": {"kind": "tensor_product", "factors": [
        {"kind": "matern", "input_shape": [], "nu": 2.5, "lengthscales": ell},
        {"kind": "matern", "input_shape": [], "nu": 2.5, "lengthscales": ell}]}}
    kf_spec = {"scale": 100.0, "base": {"kind": "expquad", "input_shape": [2], "lengthscales": [0.3, 0.4]}}
    k_u, k_f = helpers.api_kernel(ku_spec), helpers.api_kernel(kf_spec)
    u_prior = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), k_u)
    f_prior = lg.GaussianProcess(lg.functions.Constant(input_shape=(2,), value=1.5), k_f)
    Xb = np.stack([np.linspace(0, 1, 33), np.zeros(33)], -1)
    Xp = rng.uniform(0, 1, (301, 2))
    Xt = rng.uniform(0, 1, (57, 2))
    L = -1.0 * diffops.Laplacian((2,))
    Yb, Yp = np.zeros(33), np.zeros(301)
    b_lazy = -f_prior(Xp)
    assert isinstance(b_lazy.cov, lg.linops.CovarianceLinearOperator)
    b_dense = lg.randvars.Normal(b_lazy.mean, b_lazy.dense_cov)
    posts = {}
    for name, b in (("lazy", b_lazy), ("dense", b_dense)):
        posts[name] = u_prior.condition_on_observations(Yb, X=Xb).condition_on_observations(Yp, X=Xp, L=L, b=b)
    posts["oneshot"] = lg.ConditionalGaussianProcess.from_observation_batches(u_prior, [(Yb, Xb), (Yp, Xp, L, b_lazy)])
    posts["blockrows"] = lg.ConditionalGaussianProcess.from_observation_batches(
        u_prior, [(Yb, Xb), (Yp, Xp, L, b_lazy)], nb=128, replicate=False)
    # numpy formula from the oracle's matrices
    lap = [(-1.0, ("wl", np.ones(2)))]
    G = np.block([[ocf.matrix(ku_spec, None, None, Xb), ocf.matrix(ku_spec, None, lap, Xb, Xp)],
                  [ocf.matrix(ku_spec, lap, None, Xp, Xb), ocf.matrix(ku_spec, lap, lap, Xp) + ocf.matrix(kf_spec, None, None, Xp)]])
    KtX = np.hstack([ocf.matrix(ku_spec, None, None, Xt, Xb), ocf.matrix(ku_spec, None, lap, Xt, Xp)])
    y = np.concatenate([Yb, Yp - (-1.5)])  # residual: Y - (L m_u + b.mean), b.mean = -1.5
    mean_ref = KtX @ np.linalg.solve(G, y)
    var_ref = 4.0 - np.einsum("ij,ji->i", KtX, np.linalg.solve(G, KtX.T))
    sc = max(np.max(np.abs(mean_ref)), 4.0)
    for name, post in posts.items():
        assert np.max(np.abs(post.mean(Xt) - mean_ref)) <= 1e-8 * sc, name
        assert np.max(np.abs(post.var(Xt) - var_ref)) <= 1e-8 * sc, name


# ---- tensor-grid (Kronecker) structure path: SURVEY.md section 8f item 3 ------------------------------------------
_KR = np.load(os.path.join(GOLDEN, "kron.npz"))
KRON_SPECS = json.loads(bytes(_KR["__specs__"]).decode())


@pytest.mark.parametrize("spec", KRON_SPECS, ids=[s["name"] for s in KRON_SPECS])
def test_kronecker_linop_on_tensor_product_grids_matches_reference(spec):
    """``k.linop(TensorProductGrid, TensorProductGrid | None)`` yields (sums of) Kronecker products whose dense form
    (Kronecker assembly kernel) and structured product (two DMMA GEMMs per term) match the frozen outputs of the
    reference's Kronecker linops; the same kernel evaluated pair by pair on the flattened grid agrees too."""
    from linpde_gp_b200 import linops
    from linpde_gp_b200.randprocs import covfuncs

    k = helpers.api_L0kL1(spec)
    g0 = covfuncs.TensorProductGrid(*[np.asarray(f) for f in spec["factors0"]])
    g1 = None if spec["factors1"] is None else covfuncs.TensorProductGrid(*[np.asarray(f) for f in spec["factors1"]])
    K_ref, V, KV_ref = _KR[spec["name"] + "__K"], _KR[spec["name"] + "__V"], _KR[spec["name"] + "__KV"]
    sc = np.max(np.abs(K_ref))
    op = k.linop(g0, g1)
    assert op.kron_terms() is not None, "grid inputs must keep the Kronecker structure"
    assert isinstance(op, (linops.Kronecker, linops.ScaledLinearOperator, linops.SumLinearOperator))
    assert op.shape == K_ref.shape
    K = op.todense()
    assert np.max(np.abs(K - K_ref)) <= GRAM_TOL * sc
    assert np.max(np.abs(op @ V - KV_ref)) <= 1e-11 * np.max(np.abs(KV_ref))
    assert np.max(np.abs(op @ V[:, 0] - KV_ref[:, 0])) <= 1e-11 * np.max(np.abs(KV_ref))
    # plain point sets (the grid flattened C-order) take the pairwise Gram kernel: same matrix
    d = len(spec["factors0"])
    Kp = k.matrix(np.asarray(g0).reshape(-1, d), None if g1 is None else np.asarray(g1).reshape(-1, d))
    assert np.max(np.abs(Kp - K_ref)) <= GRAM_TOL * sc
    A = K_ref + 1e-6 * sc * np.eye(len(K_ref)) if g1 is None else None
    # SPD solves straight from the Kronecker assembly (lower mode into the factor buffer).  (A fourth derivative of a
    # Matern-3/2 factor does not exist in the mean-square sense: kron_tp3_LkL is a formula check only, its matrix is
    # indefinite in the reference as well.)
    if A is not None and np.min(np.linalg.eigvalsh(A)) > 0:
        noisy = op + linops.Scaling(np.full(len(K_ref), 1e-6 * sc))  # flagged symmetric, like pn.linops.Scaling
        x = noisy.solve(V)
        assert np.max(np.abs(A @ x - V)) <= 1e-7 * np.max(np.abs(V))


@pytest.mark.parametrize("shapes", [((1, 1), (1, 1)), ((3, 5), (7, 2)), ((4, 4), (65, 65)), ((33, 17), (9, 31)), ((130, 2), (3, 129))])
def test_kron_sum_kernel_vs_torch(shapes):
    """lpgp_kron_sum against torch.kron: ragged shapes, odd leading dimensions, more terms than one launch takes,
    accumulate mode and the lower-triangle mode."""
    import torch

    from linpde_gp_b200 import backend as be

    (n1, m1), (n2, m2) = shapes
    g = torch.Generator(device="cuda").manual_seed(n1 * 1000 + m2)
    for nterms in (1, 3, 6):
        terms = []
        for t in range(nterms):
            A = be.alloc_matrix(n1, m1).normal_(generator=g)
            B = torch.randn(n2, m2 + 1, dtype=torch.float64, device="cuda", generator=g)[:, :m2]  # odd ld / unaligned rows
            terms.append((0.5 + t, A, B))
        ref = sum(a * torch.kron(A, B) for a, A, B in terms)
        out = be.kron_sum(terms)
        assert out.shape == ref.shape
        assert (out - ref).abs().max().item() <= 1e-13 * max(ref.abs().max().item(), 1.0)
        out2 = be.kron_sum(terms, out=out, accumulate=True)
        assert (out2 - 2 * ref).abs().max().item() <= 1e-13 * max(ref.abs().max().item(), 1.0)
    if n1 * n2 == m1 * m2:
        n = n1 * n2
        out = be.alloc_matrix(n, n).fill_(-7.0)
        be.kron_sum(terms, out=out, lower=True)
        low = torch.tril(out)
        assert (low - torch.tril(ref)).abs().max().item() <= 1e-13 * max(ref.abs().max().item(), 1.0)
        if n > 512:  # tiles strictly above the diagonal stay untouched
            assert (out[:64, 320:] == -7.0).all()


def test_gridded_conditioning_uses_kronecker_assembly_and_matches_pairwise():
    """Conditioning on TensorProductGrid batches assembles the Gram blocks from Kronecker factors; the posterior is
    the one obtained from the same points passed as plain arrays (pairwise Gram kernel) and the oracle's."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200._lib import lib
    from oracle import gp as ogp

    prob = ogp.poisson2d_grid_problem(nx=24, ny=20, n_bc_edge=30, ell=0.2, grid=14)
    ref = ogp.solve(prob)
    post_g, res_g = helpers.api_solve(prob)
    assert all(b.grid is not None for b in post_g._blocks)
    flat = dict(prob, blocks=[{k: v for k, v in b.items() if k != "grid"} for b in prob["blocks"]])
    post_p, res_p = helpers.api_solve(flat)
    assert all(b.grid is None for b in post_p._blocks)
    gsc = np.max(np.abs(ref["gram"]))
    assert np.max(np.abs(res_g["gram"] - ref["gram"])) <= GRAM_TOL * gsc
    assert np.max(np.abs(res_g["gram"] - res_p["gram"])) <= GRAM_TOL * gsc
    for key in ("mean", "var", "cov"):
        sc = max(np.max(np.abs(ref[key])), np.max(np.abs(ref["var"])))
        assert np.max(np.abs(res_g[key] - ref[key])) <= POST_TOL * sc, key
        assert np.max(np.abs(res_g[key] - res_p[key])) <= POST_TOL * sc, key


# ---- multi-output processes with independent outputs ------------------------------------------------
MO_FILES = sorted(glob.glob(os.path.join(GOLDEN, "mo_*.npz")))


@pytest.mark.parametrize("one_shot", [False, True], ids=["sequential", "one_shot"])
@pytest.mark.parametrize("path", MO_FILES, ids=[os.path.basename(p)[3:-4] for p in MO_FILES])
def test_multi_output_conditioning_matches_reference_golden(path, one_shot):
    """IndependentMultiOutputCovarianceFunction prior, ``D @ SelectOutput(i) - SelectOutput(j)`` observations
    (experiments/0000_cpu_stationary_1d.ipynb cells 55-82), ``SelectOutput(j)(posterior)`` mean / variance /
    covariance against the frozen outputs of the real reference; Gram 1e-11 (reconstructed as L L^T), posterior 1e-8."""
    g = np.load(path)
    problem = json.loads(bytes(g["problem"]).decode())
    if one_shot and any("functional" in blk for blk in problem["blocks"]):
        # integral / sum functionals (cpu_1d_stat*: the stationarity condition of notebook cells 65-66, 85) are
        # appended with condition_on_observations; the one-shot path takes plain L[f](X) batches only
        with pytest.raises(NotImplementedError):
            helpers.api_solve_multi_output(problem, one_shot=True)
        return
    post, res = helpers.api_solve_multi_output(problem, one_shot=one_shot)
    gs = np.max(np.abs(g["gram"]))
    assert np.max(np.abs(res["gram"] - g["gram"])) <= 1e-11 * gs
    for key in ("mean", "var", "cov"):
        for j in range(len(problem["kernels"])):
            sc = max(np.max(np.abs(g[key][j])), np.max(np.abs(g["var"][j])))
            assert np.max(np.abs(res[key][j] - g[key][j])) <= POST_TOL * sc, (key, j)
    assert np.max(np.abs(res["w"] - g["w"])) <= 1e-6 * np.max(np.abs(g["w"]))
    # the un-selected posterior: mean of all outputs at once, variance via the per-output posteriors
    Xt = np.asarray(problem["Xt"], dtype=float)
    m_all = post.mean(Xt)
    assert m_all.shape == Xt.shape[:1] + (len(problem["kernels"]),)
    np.testing.assert_allclose(m_all.T, res["mean"], rtol=0, atol=1e-12 * np.max(np.abs(res["mean"])))
    with pytest.raises(NotImplementedError):
        post.cov(Xt, None)


_I = np.load(os.path.join(GOLDEN, "integrals.npz"))


def test_matern_lebesgue_integral_kernels_match_reference_golden():
    """``lpgp_matern_integral`` / ``lpgp_matern_integral2`` against the frozen outputs of the reference's
    ``UnivariateHalfIntegerMaternLebesgueIntegral`` on its own test cases (cases_integral_matern.py), 1e-12 of max;
    strided / accumulating / weighted output modes against the same numbers."""
    import torch

    from linpde_gp_b200 import _lowering, backend

    worst = 0.0
    for i, nu in enumerate(_I["nus"]):
        for j, ell in enumerate(_I["lengthscales"]):
            dsc = _lowering.matern_integral_desc(float(nu), float(ell))
            for d, (a, b) in enumerate(_I["domains"]):
                x = backend.to_device(_I["X"][d])
                ref = _I["Lk"][i, j, d]
                out = torch.empty(10, dtype=torch.float64, device=x.device)
                backend.matern_integral(dsc, a, b, x, out)
                worst = max(worst, np.max(np.abs(out.cpu().numpy() - ref)) / np.max(np.abs(ref)))
                # column of a row-major matrix, accumulated on top of existing content, scaled by alpha and a device weight
                K = torch.ones((10, 16), dtype=torch.float64, device=x.device)
                w = torch.full((1,), -2.0, dtype=torch.float64, device=x.device)
                backend.matern_integral(dsc, a, b, x, K[:, 3:], out_stride=16, alpha=0.5, w=w, accumulate=True)
                np.testing.assert_allclose(K[:, 3].cpu().numpy(), 1.0 - ref, rtol=0, atol=1e-12 * np.max(np.abs(ref)))
                assert float(K.sum()) == pytest.approx(160.0 - ref.sum(), abs=1e-9)  # nothing else touched
            for d, (d0, d1) in enumerate(_I["domain_pairs"]):
                out = torch.zeros(1, dtype=torch.float64, device="cuda")
                backend.matern_integral2(dsc, tuple(d0), tuple(d1), out)
                backend.matern_integral2(dsc, tuple(d0), tuple(d1), out, alpha=2.0, accumulate=True)
                ref = _I["LkL"][i, j, d]
                worst = max(worst, abs(float(out.item()) / 3.0 - ref) / abs(ref))
    assert worst <= 1e-12, worst
    backend.matern_integral(dsc, 0.0, 1.0, torch.empty(0, dtype=torch.float64, device="cuda"), out)  # empty input: no-op
    with pytest.raises(ValueError):
        backend.matern_integral(dsc, 0.0, 1.0, x, out, out_stride=0)


def test_integral_observation_of_scalar_process():
    """``LebesgueIntegral`` observation of a SCALAR process with a sum kernel (two closed-form terms per entry) next to
    point observations; checked against the numpy formula built from the oracle's closed forms."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.randprocs import covfuncs
    from oracle import covfuncs as ocf
    from oracle import integrals as oint
    from oracle import linalg as ola

    s1 = {"scale": 2.0, "base": {"kind": "matern", "input_shape": [], "nu": 2.5, "lengthscales": 0.3}}
    s2 = {"scale": 0.5, "base": {"kind": "matern", "input_shape": [], "nu": 0.5, "lengthscales": 0.7}}
    k = 2.0 * covfuncs.Matern((), nu=2.5, lengthscales=0.3) + 0.5 * covfuncs.Matern((), nu=0.5, lengthscales=0.7)
    rng = np.random.default_rng(3)
    X, Xt = rng.uniform(0, 1, 12), np.linspace(-0.2, 1.2, 29)
    Y = np.cos(2 * X)
    dom = (0.1, 0.9)
    prior = lg.GaussianProcess(lg.functions.Constant((), 0.5), k)
    post = prior.condition_on_observations(Y, X=X).condition_on_observations(0.3, L=1.5 * lg.linfunctls.LebesgueIntegral(dom))
    kk = lambda a, b=None: ocf.matrix(s1, None, None, a, b) + ocf.matrix(s2, None, None, a, b)  # noqa: E731
    ki = lambda x: 1.5 * (oint.integral_crosscov(s1, dom, x) + oint.integral_crosscov(s2, dom, x))  # noqa: E731
    G = np.block([[kk(X), ki(X)[:, None]], [ki(X)[None, :], np.full((1, 1), 2.25 * (
        oint.integral_integral(s1, dom, dom) + oint.integral_integral(s2, dom, dom)))]])
    resid = np.concatenate([Y - 0.5, [0.3 - 1.5 * 0.5 * 0.8]])
    Lc = ola.cholesky_lower(G)
    Kt = np.concatenate([kk(Xt, X), ki(Xt)[:, None]], axis=1)
    mean = 0.5 + Kt @ ola.cho_solve_lower(Lc, resid)
    cov = kk(Xt) - Kt @ ola.cho_solve_lower(Lc, Kt.T)
    np.testing.assert_allclose(post.gram.todense(), G, rtol=0, atol=1e-11 * np.max(np.abs(G)))
    np.testing.assert_allclose(post.mean(Xt), mean, rtol=0, atol=POST_TOL * np.max(np.abs(mean)))
    np.testing.assert_allclose(post.var(Xt), np.diag(cov), rtol=0, atol=POST_TOL * 2.5)
    np.testing.assert_allclose(post.cov.matrix(Xt), cov, rtol=0, atol=POST_TOL * 2.5)
    # an operator on the integrated kernel has no closed form (neither in the reference): loud failure, no fallback
    with pytest.raises(NotImplementedError):
        post.condition_on_observations(np.zeros(3), X=np.linspace(0, 1, 3), L=lg.linfuncops.diffops.Laplacian(()))
    with pytest.raises(NotImplementedError):
        lg.GaussianProcess(lg.functions.Zero(()), covfuncs.ExpQuad(())).condition_on_observations(
            0.0, L=lg.linfunctls.LebesgueIntegral(dom))


def test_multi_output_prior_kernel_evaluation():
    """tests/linpde_gp/randprocs/kernels/test_independent_multi_output.py: independence, batched shapes,
    block-diagonal ``linop`` with non-square blocks."""
    from linpde_gp_b200.randprocs import covfuncs

    ls = np.random.default_rng(12938422).random(size=(3, 2))
    ks = [covfuncs.TensorProduct(*(covfuncs.Matern((), nu=2.5, lengthscales=l) for l in row)) for row in ls]
    mo = covfuncs.IndependentMultiOutputCovarianceFunction(*ks)
    rng = np.random.default_rng(9238134)
    x0, x1 = rng.random(size=(10, 1, 2)), rng.random(size=(1, 15, 2))
    res = mo(x0, x1)
    assert res.shape == (10, 15, 3, 3)
    for i, j in np.ndindex(3, 3):
        if i != j:
            assert np.all(res[..., i, j] == 0.0)
        else:
            np.testing.assert_allclose(res[..., i, i], ks[i](x0, x1), rtol=0, atol=1e-15)
    same = mo(x0[:, 0], x0[:, 0])
    np.testing.assert_allclose(same[..., np.arange(3), np.arange(3)], 1.0, atol=1e-15)
    dense = mo.linop(x0, x1).todense()
    assert dense.shape == (30, 45)
    for i in range(3):
        np.testing.assert_allclose(dense[10 * i : 10 * i + 10, 15 * i : 15 * i + 15], ks[i].matrix(x0, x1), atol=1e-15)
    assert np.count_nonzero(dense[:10, 15:]) == 0
    st = covfuncs.StackCovarianceFunction((ks[0], covfuncs.Zero((2,)), ks[2]), output_idx=0)
    assert st(x0, x1).shape == (10, 15, 3) and st.linop(x0, x1).shape == (30, 15)
    np.testing.assert_allclose(st.linop(x0, x1).todense()[20:], ks[2].matrix(x0, x1), atol=1e-15)


def test_sum_kernel_prior_posterior_covariance():
    """Posterior covariance with a SUM prior kernel of different base factors (two descriptors per block): the
    cross-covariance workspace accumulates consecutive entries on the same columns."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.randprocs import covfuncs
    from oracle import covfuncs as ocf
    from oracle import linalg as ola

    rng = np.random.default_rng(5)
    X, Xt = rng.uniform(0, 1, 40), np.linspace(0, 1, 23)
    Y = np.sin(3 * X)
    k = 2.0 * covfuncs.Matern((), nu=2.5, lengthscales=0.3) + 0.5 * covfuncs.ExpQuad((), lengthscales=0.2)
    post = lg.GaussianProcess(lg.functions.Zero(()), k).condition_on_observations(
        Y, X=X, b=lg.randvars.Normal(np.zeros(40), lg.linops.Scaling(1e-4 * np.ones(40))))
    s1 = {"scale": 2.0, "base": {"kind": "matern", "input_shape": [], "nu": 2.5, "lengthscales": 0.3}}
    s2 = {"scale": 0.5, "base": {"kind": "expquad", "input_shape": [], "lengthscales": 0.2}}
    kk = lambda a, b=None: ocf.matrix(s1, None, None, a, b) + ocf.matrix(s2, None, None, a, b)
    G = kk(X) + 1e-4 * np.eye(40)
    Lc = ola.cholesky_lower(G)
    Kt = kk(Xt, X)
    mean = Kt @ ola.cho_solve_lower(Lc, Y)
    cov = kk(Xt) - Kt @ ola.cho_solve_lower(Lc, Kt.T)
    np.testing.assert_allclose(post.mean(Xt), mean, rtol=0, atol=POST_TOL * np.max(np.abs(mean)))
    np.testing.assert_allclose(post.cov.matrix(Xt), cov, rtol=0, atol=POST_TOL * 2.5)
    np.testing.assert_allclose(post.var(Xt), np.diag(cov), rtol=0, atol=POST_TOL * 2.5)


def test_heat_ibvp_analytic_solution_within_two_sigma():
    """The reference's end-to-end test, tests/linpde_gp/problems/test_heat.py:56-99, restated: 1-D heat equation on
    t in [0, 5], x in [-1, 1], alpha = 0.1, initial values 1 sin(w1 (x+1)) + 2 sin(w2 (x+1)) (TruncatedSineSeries
    [1, 2], w_n = n pi / 2), prior Matern-3/2(t; 2.5) x Matern-5/2(x; 2.0); N_ic = 5 (inset 1e-6), N_bc = 2 x 50 with
    noise 1e-5, N_pde = 100 x 20 uniform grid.  Same assertions and tolerances as the reference: observations are
    reproduced to atol 3e-2 and the analytic solution (problems/pde/_heat.py:96-131) lies within mean +- 2 std
    (slack 3e-2) on a 50 x 50 grid."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.linfuncops import diffops
    from linpde_gp_b200.randprocs import covfuncs

    t0, T, lo, hi, alpha = 0.0, 5.0, -1.0, 1.0, 0.1
    coeffs = np.array([1.0, 2.0])
    omega = np.arange(1, 3) * np.pi / (hi - lo)

    def solution(tx):
        t, x = tx[..., :1], tx[..., 1:]
        return np.sum(coeffs * np.sin(omega * (x - lo)) * np.exp(alpha * omega**2 * (t0 - t)), axis=-1)

    def grid(ts, xs):
        return np.stack(np.meshgrid(ts, xs, indexing="ij"), axis=-1)

    def noise(X):
        n = int(np.prod(X.shape[:-1]))
        return lg.randvars.Normal(np.zeros(X.shape[:-1]), np.diag(1e-5 * np.ones(n)))

    def assert_observations_match(X, Y, gp, tol=3e-2):
        assert np.allclose(gp.mean(X), Y, rtol=0.0, atol=tol)

    prior = lg.GaussianProcess(
        lg.functions.Zero(input_shape=(2,)),
        1.0**2 * covfuncs.TensorProduct(covfuncs.Matern((), nu=1.5, lengthscales=2.5),
                                        covfuncs.Matern((), nu=2.5, lengthscales=2.0)))
    X_ic = grid(np.array([t0]), np.linspace(lo + 1e-6, hi - 1e-6, 5))[0]
    Y_ic = solution(X_ic)
    u = prior.condition_on_observations(Y_ic, X_ic)
    assert_observations_match(X_ic, Y_ic, u)
    for xb in (lo, hi):
        X_bc = grid(np.linspace(t0, T, 50), np.array([xb]))[:, 0]
        Y_bc = np.zeros(50)
        u = u.condition_on_observations(Y_bc, X=X_bc, b=noise(X_bc))
        assert_observations_match(X_bc, Y_bc, u)
    X_pde = grid(np.linspace(t0, T, 100), np.linspace(lo, hi, 20))
    u = u.condition_on_observations(np.zeros((100, 20)), X=X_pde, L=diffops.HeatOperator((2,), alpha=alpha))
    X_test = grid(np.linspace(t0, T, 50), np.linspace(lo, hi, 50))
    Y_test = solution(X_test)
    mean, std = u.mean(X_test), np.nan_to_num(u.std(X_test))
    assert mean.shape == (50, 50) and std.shape == (50, 50)
    assert np.min(mean + 2 * std - Y_test) > -3e-2
    assert np.min(Y_test - (mean - 2 * std)) > -3e-2
    assert np.max(np.abs(mean - Y_test)) < 0.1  # and the mean itself is a decent solution of the IBVP


def test_polynomial_prior_mean_is_pushed_through_operators_in_closed_form():
    """A polynomial prior mean m: conditioning f = m + g on ``L f (X) = Y`` equals conditioning the zero-mean g on
    ``Y - (L m)(X)`` and adding m back (_conditional.py:193-197, 296-399) -- ``L m`` by exact coefficient calculus
    (functions/_polynomial.py:88-96) instead of the reference's JAX fallback.  1e-10 relative."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.functions import Polynomial
    from linpde_gp_b200.linfuncops import diffops
    from linpde_gp_b200.randprocs import covfuncs

    m = Polynomial([0.5, -1.0, 0.25, 2.0])
    k = 1.5 * covfuncs.Matern((), nu=2.5, lengthscales=0.7)
    L = -1.0 * diffops.Laplacian(())
    X_pde, X_bc = np.linspace(-0.8, 0.8, 17), np.array([-1.0, 1.0])
    Y_pde, Y_bc = np.pi**2 * np.sin(np.pi * X_pde), np.zeros(2)
    xs = np.linspace(-1, 1, 41)

    post = lg.GaussianProcess(m, k).condition_on_observations(Y_bc, X=X_bc).condition_on_observations(Y_pde, X=X_pde, L=L)
    Lm = L(m)
    np.testing.assert_allclose(Lm(X_pde), -(0.5 + 12.0 * X_pde), rtol=1e-14)
    post0 = (lg.GaussianProcess(lg.functions.Zero(()), k)
             .condition_on_observations(Y_bc - m(X_bc), X=X_bc)
             .condition_on_observations(Y_pde - Lm(X_pde), X=X_pde, L=L))
    sc = np.max(np.abs(post.mean(xs)))
    assert np.max(np.abs(post.mean(xs) - (m(xs) + post0.mean(xs)))) <= 1e-10 * sc
    assert np.max(np.abs(post.var(xs) - post0.var(xs))) <= 1e-10 * 1.5
    assert np.max(np.abs(post.mean(X_bc) - Y_bc)) <= 1e-8 * sc
    # the push-forward L(posterior) carries the mean too
    res = L(post).mean(X_pde)
    assert np.max(np.abs(res - Y_pde)) <= 1e-7 * np.max(np.abs(Y_pde))


def test_reference_heat_test_verbatim_through_problem_classes():
    """tests/linpde_gp/problems/test_heat.py:9-99 of the reference with only the package name changed: problem
    definition through ``problems.pde.HeatEquationDirichletProblem``, grids through ``domains`` (TensorProductGrids, so
    the PDE batch takes the Kronecker assembly path), same assertions and tolerances."""
    import linpde_gp_b200 as linpde_gp

    spatial_domain = linpde_gp.domains.asdomain([-1.0, 1.0])
    ibvp = linpde_gp.problems.pde.HeatEquationDirichletProblem(
        t0=0.0, T=5.0, spatial_domain=spatial_domain, alpha=0.1,
        initial_values=linpde_gp.functions.TruncatedSineSeries(spatial_domain, coefficients=[1.0, 2.0]))

    def assert_observations_match(obs, gp, tol=3e-2):
        X_obs, Y_obs = obs
        assert np.allclose(gp.mean(X_obs), Y_obs, rtol=0.0, atol=tol)

    def assert_within_uncertainty_region(obs, gp):
        X_obs, Y_obs = obs
        vals_gp, std_gp = gp.mean(X_obs), np.nan_to_num(gp.std(X_obs))
        assert np.min(vals_gp + 2 * std_gp - Y_obs) > -3e-2
        assert np.min(Y_obs - (vals_gp - 2 * std_gp)) > -3e-2

    def get_noise(X):
        num_entries = int(np.prod(X.shape[:-1]))
        return linpde_gp.randvars.Normal(np.zeros(X.shape[:-1]), np.diag(1e-5 * np.ones(num_entries)))

    u_prior = linpde_gp.GaussianProcess(
        mean=linpde_gp.functions.Zero(input_shape=(2,)),
        cov=1.0**2 * linpde_gp.randprocs.covfuncs.TensorProduct(
            linpde_gp.randprocs.covfuncs.Matern((), nu=1.5, lengthscales=2.5),
            linpde_gp.randprocs.covfuncs.Matern((), nu=2.5, lengthscales=2.0)))

    X_ic = ibvp.initial_domain.uniform_grid(5, inset=1e-6)
    Y_ic = ibvp.initial_condition.values(X_ic[..., 1])
    u_ic = u_prior.condition_on_observations(Y_ic, X_ic)
    assert_observations_match((X_ic, Y_ic), u_ic)
    u_ic_bc = u_ic
    for bc in ibvp.boundary_conditions:
        X_bc = bc.boundary.uniform_grid(50)
        Y_bc = bc.values(X_bc)
        u_ic_bc = u_ic_bc.condition_on_observations(Y_bc, X=X_bc, b=get_noise(X_bc))
        assert_observations_match((X_bc, Y_bc), u_ic_bc)
    X_pde = ibvp.domain.uniform_grid((100, 20))
    Y_pde = ibvp.pde.rhs(X_pde)
    u_ic_bc_pde = u_ic_bc.condition_on_observations(Y_pde, X=X_pde, L=ibvp.pde.diffop)
    X_test = ibvp.domain.uniform_grid((50, 50))
    Y_test = ibvp.solution(X_test)
    assert_within_uncertainty_region((X_test, Y_test), u_ic_bc_pde)


@pytest.mark.parametrize("prior_kind", ["expquad", "matern"])
def test_experiment_0000_poisson_dirichlet_1d_flow(prior_kind):
    """BASELINE.json configs[0] (experiments/0000_poisson_dirichlet_1d.ipynb): 1-D Poisson problem on [-1, 1] with constant
    right-hand side through ``problems.pde.PoissonEquationDirichletProblem``, boundary observations from
    ``get_1d_dirichlet_boundary_observations``, collocation points ``linspace(-0.8, 0.8, n)`` with n = 3 (ExpQuad) / 100
    (Matern-5/2) as in the notebook; the posterior mean must agree with the analytic solution within 2 std (+1e-6) and to
    0.1 absolutely (numpy evaluation of the same posterior: 0.066 / 0.037), and the PDE residual posterior must
    reproduce the right-hand side at the collocation points."""
    import linpde_gp_b200 as linpde_gp
    from linpde_gp_b200.randprocs import covfuncs

    domain = linpde_gp.domains.asdomain([-1.0, 1.0])
    bvp = linpde_gp.problems.pde.PoissonEquationDirichletProblem(
        domain, rhs=linpde_gp.functions.Constant(input_shape=(), value=2.0), boundary_values=(0.0, 0.0))
    cov = 2.0**2 * (covfuncs.ExpQuad((), lengthscales=1.0) if prior_kind == "expquad"
                    else covfuncs.Matern((), nu=2.5, lengthscales=1.0))
    u_prior = linpde_gp.GaussianProcess(linpde_gp.functions.Zero(input_shape=()), cov)
    X_bc, Y_bc = linpde_gp.problems.pde.get_1d_dirichlet_boundary_observations(bvp.boundary_conditions)
    u_bc = u_prior.condition_on_observations(Y_bc, X=X_bc)
    n_pde = 100 if prior_kind == "matern" else 3
    X_pde = domain.uniform_grid(n_pde, inset=0.2)
    u_post = u_bc.condition_on_observations(bvp.pde.rhs(X_pde), X=X_pde, L=bvp.pde.diffop)
    xs = domain.uniform_grid(100)
    mean, std = u_post.mean(xs), np.nan_to_num(u_post.std(xs))
    truth = bvp.solution(xs)
    assert np.max(np.abs(mean - truth)) <= 0.1
    assert np.all(np.abs(mean - truth) <= 2 * std + 1e-6)
    residual = bvp.pde.diffop(u_post)
    assert np.max(np.abs(residual.mean(X_pde) - 2.0)) <= 1e-6
    assert np.max(np.abs(u_post.mean(X_bc))) <= 1e-8


def test_experiment_0001_poisson_dirichlet_2d_flow():
    """The paper's 2-D Poisson example (experiments/0001_poisson_dirichlet_2d.ipynb; BASELINE.json configs[1] at notebook
    size): -Laplace u = 2 on [-1, 1]^2 with zero Dirichlet values, product Matern-5/2 prior, 4 x 20 boundary points and a
    20 x 20 collocation grid from ``domains`` (TensorProductGrids -> Kronecker assembly).  Checked against the classical
    series solution of the torsion problem, u = 1 - x^2 - 32/pi^3 sum_n (-1)^n/(2n+1)^3 cosh(k_n y)/cosh(k_n) cos(k_n x),
    k_n = (2n+1) pi/2: inside mean +- 2 std (+1e-3) on a 40 x 40 grid, and the posterior PDE residual reproduces the
    right-hand side at the collocation points."""
    import linpde_gp_b200 as linpde_gp
    from linpde_gp_b200.randprocs import covfuncs

    def torsion(tx, terms=60):
        x, y = tx[..., 0], tx[..., 1]
        n = np.arange(terms)
        kn = (2 * n + 1) * np.pi / 2
        s = np.sum(((-1.0) ** n / (2 * n + 1) ** 3) * np.cosh(kn * y[..., None]) / np.cosh(kn) * np.cos(kn * x[..., None]), axis=-1)
        return 1 - x**2 - 32 / np.pi**3 * s

    domain = linpde_gp.domains.asdomain([np.array([-1.0, -1.0]), np.array([1.0, 1.0])])
    bvp = linpde_gp.problems.pde.PoissonEquationDirichletProblem(domain, rhs=linpde_gp.functions.Constant((2,), 2.0))
    u = linpde_gp.GaussianProcess(
        linpde_gp.functions.Zero(input_shape=(2,)),
        2.0**2 * covfuncs.TensorProduct(covfuncs.Matern((), nu=2.5, lengthscales=1.0),
                                        covfuncs.Matern((), nu=2.5, lengthscales=1.0)))
    for bc in bvp.boundary_conditions:
        X_bc = bc.boundary.uniform_grid(20, inset=0.02)  # (without the inset the four corners are observed twice: singular)
        u = u.condition_on_observations(bc.values(X_bc), X=X_bc)
    X_pde = domain.uniform_grid((20, 20), inset=0.05)
    u = u.condition_on_observations(bvp.pde.rhs(X_pde), X=X_pde, L=bvp.pde.diffop)
    X_test = domain.uniform_grid((40, 40))
    truth = torsion(np.asarray(X_test))
    mean, std = u.mean(X_test), np.nan_to_num(u.std(X_test))
    assert mean.shape == (40, 40)
    # numpy evaluation of the same posterior (cond(G) = 1.1e9): max error 0.144 (at the unobserved corners), always
    # 2.4e-3 inside the 2-std band
    assert np.max(np.abs(mean - truth)) <= 0.2
    assert np.all(np.abs(mean - truth) <= 2 * std + 1e-3)
    res = bvp.pde.diffop(u).mean(X_pde)
    assert np.max(np.abs(res - 2.0)) <= 1e-5
    assert np.max(np.abs(mean - mean.T)) <= 1e-5 and np.max(np.abs(mean - mean[::-1, :])) <= 1e-5  # symmetries of the problem


# ---- appendable factor: capacity reserve, versioned extent, no reference cycles -----------------
def _poisson_prior_and_batches(n_bc=64, n_pde=600, seed=11):
    import linpde_gp_b200 as lg
    from linpde_gp_b200.linfuncops import diffops
    from linpde_gp_b200.randprocs import covfuncs

    rng = np.random.default_rng(seed)
    ell = 0.25
    k = 4.0 * covfuncs.TensorProduct(covfuncs.Matern((), nu=2.5, lengthscales=ell), covfuncs.Matern((), nu=2.5, lengthscales=ell))
    prior = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), k)
    s = np.linspace(0, 1, n_bc, endpoint=False)
    Xb = np.stack([s, np.zeros_like(s)], -1)
    Xp = rng.uniform(0.05, 0.95, (n_pde, 2))
    Xq = rng.uniform(0.05, 0.95, (130, 2))
    lap = -1.0 * diffops.Laplacian((2,))
    return prior, (np.zeros(n_bc), Xb, None), (np.full(n_pde, 2.0), Xp, lap), (np.sin(Xq[:, 0]), Xq, None)


def test_append_in_place_shares_storage_and_keeps_old_posterior_valid():
    """extended() allocates and copies nothing while the reserved capacity suffices; the posterior conditioned on fewer
    batches keeps evaluating to the same numbers after its factor buffer has grown in place."""
    prior, b0, b1, b2 = _poisson_prior_and_batches()
    Xt = np.random.default_rng(0).uniform(0, 1, (200, 2))
    p0 = prior.condition_on_observations(b0[0], X=b0[1])
    p1 = p0.condition_on_observations(b1[0], X=b1[1], L=b1[2])
    m1, v1 = p1.mean(Xt), p1.var(Xt)
    w1 = p1.representer_weights.copy()
    f1 = p1._factor
    assert f1._storage is p0._factor._storage  # 64 + 600 rows fit the reserve (>= 1024 rows) of the first factor
    torch.cuda.synchronize()
    before = torch.cuda.memory_allocated()
    f2 = f1.extended(130)
    torch.cuda.synchronize()
    assert torch.cuda.memory_allocated() == before  # no allocation, no copy
    assert f2._storage is f1._storage and f2.L.data_ptr() == f1.L.data_ptr()
    p2 = p1.condition_on_observations(b2[0], X=b2[1])
    # p1's factor is no longer the tip of the storage (f2 claimed rows): p2 had to branch into its own storage
    assert p2._factor._storage is not f1._storage
    assert np.array_equal(p1.mean(Xt), m1) and np.array_equal(p1.var(Xt), v1)
    assert np.array_equal(p1.representer_weights, w1)
    # both branches agree with one-shot conditioning on all three batches
    import linpde_gp_b200 as lg

    ref = lg.ConditionalGaussianProcess.from_observation_batches(prior, [b0, b1, b2])
    sc = max(np.max(np.abs(ref.mean(Xt))), 4.0)
    assert np.max(np.abs(p2.mean(Xt) - ref.mean(Xt))) <= 1e-8 * sc
    assert np.max(np.abs(p2.var(Xt) - ref.var(Xt))) <= 1e-8 * sc


def test_append_beyond_capacity_moves_only_the_lower_triangle():
    from linpde_gp_b200 import backend

    rng = np.random.default_rng(3)
    n, extra = 384, 130
    A = rng.standard_normal((n + extra, n + extra))
    G = A @ A.T / (n + extra) + np.eye(n + extra)
    f = backend.DeviceFactor([n], reserve_rows=0)
    f.L.copy_(backend.to_device(G[:n, :n]))
    f.potrf()
    g = f.extended(extra)  # no capacity: new storage
    assert g._storage is not f._storage and g.capacity >= n + extra
    g.L[n:, :].copy_(backend.to_device(G[n:, :]))
    g.append_last()
    L = torch.tril(g.L).cpu().numpy()
    assert np.max(np.abs(L @ L.T - G)) <= 1e-12 * np.max(np.abs(G))
    # the old factor is untouched
    L0 = torch.tril(f.L).cpu().numpy()
    assert np.max(np.abs(L0 @ L0.T - G[:n, :n])) <= 1e-12 * np.max(np.abs(G))


def test_posterior_objects_hold_no_reference_cycles():
    """The factor of a dropped posterior is released by reference counting (cyclic GC disabled): dead 34 GB factors must
    not pile up between conditioning steps."""
    import gc
    import weakref

    prior, b0, b1, _ = _poisson_prior_and_batches()
    gc.collect()
    gc.disable()
    try:
        post = prior.condition_on_observations(b0[0], X=b0[1])
        post.mean(b0[1])
        ref = weakref.ref(post._factor._storage)
        del post
        assert ref() is None
    finally:
        gc.enable()


# ---- parity AT SIZE: frozen outputs of the real reference at N = 4,096, the oracle at 16,384 ----
def _large_golden(name):
    from oracle import make_golden_large as mgl

    path = os.path.join(GOLDEN, f"large_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    z = np.load(path)
    spec = json.loads(bytes(z["problem_spec"]).decode())
    return mgl.large_problem(spec), z


@pytest.mark.parametrize("name", ["c2_4096", "c3_4096", "c2_16384"])
def test_posterior_at_size_matches_frozen_reference(name):
    """Scaled-down configs 2 / 3 at N = 4,096 against outputs of the REAL reference (oracle/make_golden_large.py), and
    config 2 at its full N = 16,384 against the (reference-pinned) oracle: mean / variance / covariance within 1e-8."""
    prob, z = _large_golden(name)
    post, res = _api_solve_no_gram(prob)
    sc = max(np.max(np.abs(z["mean"])), np.max(np.abs(z["var"])))
    for key in ("mean", "var", "cov"):
        assert np.max(np.abs(res[key] - z[key])) <= POST_TOL * sc, (name, key, np.max(np.abs(res[key] - z[key])) / sc)
    # representer weights: compared through the residual G w = y of the reference's weights (w itself is sensitive to the
    # conditioning of the heat problem: the oracle and the reference already differ by 7e-7 there)
    w_ref = z["w"]
    assert res["w"].shape == w_ref.shape
    tol_w = 1e-8 if name.startswith("c2") else 1e-4
    assert np.max(np.abs(res["w"] - w_ref)) <= tol_w * np.max(np.abs(w_ref)), name


def _api_solve_no_gram(problem):
    """helpers.api_solve without the dense Gram read-back (2.1 GB at N = 16,384)."""
    import linpde_gp_b200 as lg

    kernel = problem["kernel"]
    shape = gcases.kernel_input_shape(kernel)
    post = lg.GaussianProcess(lg.functions.Zero(input_shape=shape), helpers.api_kernel(kernel))
    for blk in problem["blocks"]:
        X, Y = np.asarray(blk["X"], dtype=float), np.asarray(blk["Y"], dtype=float)
        b = None
        if blk.get("noise_var") is not None:
            nv = np.broadcast_to(np.asarray(blk["noise_var"], dtype=float), Y.shape).copy()
            b = lg.randvars.Normal(np.zeros_like(Y), lg.linops.Scaling(nv))
        post = post.condition_on_observations(Y, X=X, L=helpers.api_op(blk["L"]), b=b)
    Xt = np.asarray(problem["Xt"], dtype=float)
    return post, {"w": post.representer_weights, "mean": post.mean(Xt), "var": post.cov(Xt, None),
                  "cov": post.cov.matrix(Xt[: problem.get("n_cov", 16)])}


def test_inverse_rhs_conditioning_at_n4096_one_shot_and_block_rows():
    """BASELINE.json configs[4]'s inverse-RHS variant at a size the CPU can check (N = 4,096): Gram =
    L k_u L*(X, X) + k_f(X, X) accumulated on the device through ``from_observation_batches`` (single-GPU one-shot and
    the block-row distributed layout), against the numpy formula built from the oracle's matrices."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.linfuncops import diffops
    from oracle import covfuncs as ocf

    rng = np.random.default_rng(5)
    n_bc, n_pde = 256, 3840
    ell = 4.0 / np.sqrt(n_pde)
    ku_spec = {"scale": 4.0, "base": {"kind": "tensor_product", "factors": [
        {"kind": "matern", "input_shape": [], "nu": 2.5, "lengthscales": ell},
        {"kind": "matern", "input_shape": [], "nu": 2.5, "lengthscales": ell}]}}
    kf_spec = {"scale": 100.0, "base": {"kind": "expquad", "input_shape": [2], "lengthscales": [0.3, 0.4]}}
    u_prior = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), helpers.api_kernel(ku_spec))
    f_prior = lg.GaussianProcess(lg.functions.Constant(input_shape=(2,), value=1.5), helpers.api_kernel(kf_spec))
    Xb = np.stack([np.linspace(0, 1, n_bc), np.zeros(n_bc)], -1)
    Xp = rng.uniform(0, 1, (n_pde, 2))
    Xt = rng.uniform(0, 1, (200, 2))
    L = -1.0 * diffops.Laplacian((2,))
    Yb, Yp = np.zeros(n_bc), np.zeros(n_pde)
    batches = [(Yb, Xb), (Yp, Xp, L, -f_prior(Xp))]
    posts = {"oneshot": lg.ConditionalGaussianProcess.from_observation_batches(u_prior, batches),
             "blockrows": lg.ConditionalGaussianProcess.from_observation_batches(u_prior, batches, nb=512, replicate=False)}
    lap = [(-1.0, ("wl", np.ones(2)))]
    G = np.block([[ocf.matrix(ku_spec, None, None, Xb), ocf.matrix(ku_spec, None, lap, Xb, Xp)],
                  [ocf.matrix(ku_spec, lap, None, Xp, Xb), ocf.matrix(ku_spec, lap, lap, Xp) + ocf.matrix(kf_spec, None, None, Xp)]])
    KtX = np.hstack([ocf.matrix(ku_spec, None, None, Xt, Xb), ocf.matrix(ku_spec, None, lap, Xt, Xp)])
    y = np.concatenate([Yb, Yp + 1.5])
    import scipy.linalg

    cf = scipy.linalg.cho_factor(G, lower=True)
    mean_ref = KtX @ scipy.linalg.cho_solve(cf, y)
    V = scipy.linalg.solve_triangular(cf[0], KtX.T, lower=True)
    var_ref = 4.0 - np.sum(V * V, axis=0)
    sc = max(np.max(np.abs(mean_ref)), 4.0)
    for name, post in posts.items():
        assert np.max(np.abs(post.mean(Xt) - mean_ref)) <= 1e-8 * sc, name
        assert np.max(np.abs(post.var(Xt) - var_ref)) <= 1e-8 * sc, name


def test_gridded_observations_of_a_product_kernel_use_the_kronecker_factor(monkeypatch):
    """SURVEY 8f item 3 beyond dense N: plain observations on an intact TensorProductGrid under a TensorProduct prior have
    the Gram matrix ``alpha K_1 (x) K_2`` (covfuncs/_tensor_product.py:64-82, pn/linops/_kronecker.py:122-140); the
    structured factor (two small Cholesky factors) gives the same posterior as the dense bordered factor -- checked at a
    size where both run -- and conditions on a 1000 x 1200 grid (N = 1.2 M; a dense FP64 Gram matrix would be 11.5 TB)."""
    import linpde_gp_b200 as lg
    from linpde_gp_b200.randprocs import _conditional, covfuncs

    k = 1.7 * covfuncs.TensorProduct(covfuncs.Matern((), nu=2.5, lengthscales=0.3), covfuncs.Matern((), nu=1.5, lengthscales=0.5))
    prior = lg.GaussianProcess(lg.functions.Constant(input_shape=(2,), value=0.2), k)
    g1, g2 = np.linspace(0.0, 1.0, 33), np.linspace(-1.0, 1.0, 41)
    grid = covfuncs.TensorProductGrid(g1, g2)
    f = lambda x: np.sin(3 * x[..., 0]) * np.cos(2 * x[..., 1]) + 0.2  # noqa: E731
    Y = f(np.asarray(grid))
    rng = np.random.default_rng(5)
    Xt = rng.uniform([0.0, -1.0], [1.0, 1.0], (300, 2))
    monkeypatch.setattr(_conditional, "STRUCTURED_MIN_N", 10**9)
    dense = prior.condition_on_observations(Y, X=grid)
    assert not getattr(dense._factor, "structured", False)
    monkeypatch.setattr(_conditional, "STRUCTURED_MIN_N", 1000)
    post = prior.condition_on_observations(Y, X=grid)
    assert post._factor.structured and post._factor.n == 33 * 41
    assert np.max(np.abs(post.representer_weights - dense.representer_weights)) <= 1e-7 * np.max(np.abs(dense.representer_weights))
    assert np.max(np.abs(post.mean(Xt) - dense.mean(Xt))) <= 1e-9
    assert np.max(np.abs(post.var(Xt) - dense.var(Xt))) <= 1e-9 * 1.7
    C, Cd = post.cov.linop(Xt[:40]).todense(), dense.cov.linop(Xt[:40]).todense()
    assert np.max(np.abs(C - Cd)) <= 1e-9 * 1.7
    C01, Cd01 = post.cov.linop(Xt[:40], Xt[40:90]).todense(), dense.cov.linop(Xt[:40], Xt[40:90]).todense()
    assert np.max(np.abs(C01 - Cd01)) <= 1e-9 * 1.7
    b = rng.standard_normal(33 * 41)
    xs, xd = post.gram.solve(b), dense.gram.solve(b)  # cond(G) ~ 1e8: forward errors of two backward-stable solves
    assert np.max(np.abs(xs - xd)) <= 1e-5 * np.max(np.abs(xd))
    assert np.max(np.abs(post.gram @ xs - b)) <= 1e-12 * 1.7 * 33 * 41 * np.max(np.abs(xs))  # residual <= eps |G| |x|
    with pytest.raises(NotImplementedError):
        post.condition_on_observations(np.zeros(3), X=rng.uniform(0, 1, (3, 2)))
    # far beyond dense size: N = 1.2 M observations, interpolation at grid nodes, variance ~ 0 there and within the prior
    G1, G2 = np.linspace(0.0, 1.0, 1000), np.linspace(-1.0, 1.0, 1200)
    big = covfuncs.TensorProductGrid(G1, G2)
    kb = 1.7 * covfuncs.TensorProduct(covfuncs.Matern((), nu=1.5, lengthscales=0.05), covfuncs.Matern((), nu=1.5, lengthscales=0.08))
    pb = lg.GaussianProcess(lg.functions.Constant(input_shape=(2,), value=0.2), kb).condition_on_observations(f(np.asarray(big)), X=big)
    assert pb._factor.structured and pb._factor.n == 1_200_000
    nodes = np.stack([G1[[0, 17, 500, 999]], G2[[3, 600, 601, 1199]]], -1)
    assert np.max(np.abs(pb.mean(nodes) - f(nodes))) <= 1e-6
    v = pb.var(np.concatenate([nodes, Xt]))
    assert np.all(np.abs(v[:4]) <= 1e-7) and np.all(v >= -1e-7) and np.all(v <= 1.7)
    assert np.max(np.abs(pb.mean(Xt) - f(Xt))) <= 1e-3  # h = 1e-3 grid of a smooth function
