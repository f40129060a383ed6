"""GPU parity tests of the raw C-ABI kernels (through ctypes) against the frozen reference outputs, the oracle
and plain torch FP64 references.  Run on the B200 box: ``pytest -m gpu``."""
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


# every reference case lowers to a device descriptor (product form, or the radial family for the isotropic
# multi-dimensional Matern kernels)
LOWERABLE = SPECS
GRAM_TOL = 1e-12  # north_star: Gram entries rel err <= 1e-12 (relative to max |G|)


@pytest.fixture(scope="module")
def be():
    from linpde_gp_b200 import backend

    return backend


@pytest.mark.parametrize("spec", LOWERABLE, ids=[s["name"] for s in LOWERABLE])
def test_gram_matches_reference_golden(be, spec):
    desc = helpers.desc_from_spec(spec)
    shape = gcases.kernel_input_shape(spec["kernel"])
    X = gcases.sobol_points(shape)
    d = desc.d
    X0, X1 = be.points(X[:32], d), be.points(X, d)
    K = be.gram(desc, X0, X1).cpu().numpy()
    K_ref = _K[spec["name"] + "__K"]
    scale = np.max(np.abs(K_ref))
    assert np.max(np.abs(K - K_ref)) <= GRAM_TOL * scale
    dg = be.gram_diag(desc, 32).cpu().numpy()
    assert np.max(np.abs(dg - _K[spec["name"] + "__diag"])) <= GRAM_TOL * scale


@pytest.mark.parametrize("n0,n1", [(1, 1), (63, 129), (64, 128), (200, 77), (513, 1025)])
def test_gram_ragged_shapes_and_modes(be, n0, n1):
    from oracle import covfuncs as ocf

    spec = next(s for s in SPECS if s["name"] == "ns_poisson2d_LkL")
    desc = helpers.desc_from_spec(spec)
    rng = np.random.default_rng(n0 * 1000 + n1)
    A, B = rng.uniform(0, 1, (n0, 2)), rng.uniform(0, 1, (n1, 2))
    L = gcases.spec_to_oracle_op(spec["L0"])
    K_ref = ocf.matrix(spec["kernel"], L, L, A, B)
    scale = np.max(np.abs(K_ref))
    K = be.gram(desc, be.points(A, 2), be.points(B, 2))
    assert np.max(np.abs(K.cpu().numpy() - K_ref)) <= GRAM_TOL * scale
    # accumulate + alpha
    K2 = be.gram(desc, be.points(A, 2), be.points(B, 2), out=K.clone(), accumulate=True, alpha=-0.5)
    assert np.max(np.abs(K2.cpu().numpy() - 0.5 * K_ref)) <= GRAM_TOL * scale
    # odd leading dimension -> scalar store path
    buf = torch.full((n0, n1 + 3), 7.0, dtype=torch.float64, device="cuda")[:, :n1]
    if (n1 + 3) % 2 == 1:
        be.gram(desc, be.points(A, 2), be.points(B, 2), out=buf)
        assert np.max(np.abs(buf.cpu().numpy() - K_ref)) <= GRAM_TOL * scale


def test_gram_lower_mode_and_symmetrize(be):
    from oracle import covfuncs as ocf

    spec = next(s for s in SPECS if s["name"] == "ns_heat_LkL")
    desc = helpers.desc_from_spec(spec)
    rng = np.random.default_rng(5)
    A = rng.uniform(-1, 1, (333, 2))
    L = gcases.spec_to_oracle_op(spec["L0"])
    K_ref = ocf.matrix(spec["kernel"], L, L, A)
    out = torch.full((333, 336), np.nan, dtype=torch.float64, device="cuda")[:, :333]
    be.gram(desc, be.points(A, 2), None, out=out, lower=True)
    be.symmetrize_lower(out)
    assert np.max(np.abs(out.cpu().numpy() - K_ref)) <= GRAM_TOL * np.max(np.abs(K_ref))


def _set_direct_exp(flag):
    from linpde_gp_b200 import _lib

    assert _lib.lib.lpgp_set_option(_lib.OPT_DIRECT_EXP, int(flag)) == 0


@pytest.mark.parametrize("ell,offset", [(0.3, 0.0), (0.016, 0.0), (0.004, 0.0), (0.016, 1000.0), (0.0008, 0.0)])
def test_separable_matern_exponentials_match_direct_evaluation(be, ell, offset):
    """The assembly / posterior-mean kernels factor exp(-s|y-x|) = a(y) b(x) per tile (kernel_eval.cuh).  Both forms
    must agree to a few ulp of the matrix scale (far inside the 1e-12 Gram tolerance), for short lengthscales
    (|s (x - c)| up to ~560), for coordinates far from the origin, and beyond the overflow guard
    (ell = 8e-4: s * range = 2800 -> the kernel falls back to the direct form, bitwise equal)."""
    from linpde_gp_b200._lowering import Factor1D, lower

    fac = [Factor1D("matern", ell, nu=2.5), Factor1D("matern", 1.7 * ell, nu=1.5)]
    heat = {(1, 0): 1.0, (0, 2): -0.1}
    lap = {(2, 0): -1.0, (0, 2): -1.0}
    rng = np.random.default_rng(7)
    A = be.points(rng.uniform(0, 1, (300, 2)) + offset, 2)
    B = be.points(rng.uniform(0, 1, (517, 2)) + offset, 2)
    w = be.to_device(rng.standard_normal(517))
    for L0, L1 in ((None, None), (None, heat), (heat, heat), (None, {(0, 2): 1.0})):
        desc = lower(fac, L0, L1, 1.3)
        try:
            _set_direct_exp(True)
            K_dir = be.gram(desc, A, B).cpu().numpy()
            m_dir = be.post_mean(be.ObsBlocks([desc], [B], [0]), w, A).cpu().numpy()
        finally:
            _set_direct_exp(False)
        K_sep = be.gram(desc, A, B).cpu().numpy()
        m_sep = be.post_mean(be.ObsBlocks([desc], [B], [0]), w, A).cpu().numpy()
        scale = np.max(np.abs(K_dir))
        if ell < 0.001:
            assert np.array_equal(K_sep, K_dir)
        # 5e-15: the direct form itself carries ~|r| ulp of argument rounding (r up to ~40 where entries matter)
        assert np.max(np.abs(K_sep - K_dir)) <= 5e-15 * scale, (ell, offset, np.max(np.abs(K_sep - K_dir)) / scale)
        assert np.max(np.abs(m_sep - m_dir)) <= 1e-13 * np.max(np.abs(m_dir))
    del lap


def test_gram_empty_input(be):
    desc = helpers.desc_from_spec(next(s for s in SPECS if s["name"] == "ns_poisson2d_k"))
    K = be.gram(desc, torch.empty((0, 2), dtype=torch.float64, device="cuda"), torch.zeros((5, 2), dtype=torch.float64, device="cuda"))
    assert K.shape == (0, 5)


@pytest.mark.parametrize("m,n,k", [(128, 128, 8), (300, 200, 77), (1, 130, 1000), (1000, 1000, 512), (257, 511, 1030)])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (-1.0, 1.0), (0.5, -2.0)])
def test_gemm_nt_vs_torch(be, m, n, k, alpha, beta):
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    A = be.alloc_matrix(m, k).normal_(generator=g)
    B = be.alloc_matrix(n, k).normal_(generator=g)
    C = be.alloc_matrix(m, n).normal_(generator=g)
    ref = beta * C + alpha * (A @ B.T)
    be.gemm_nt(A, B, C, alpha, beta)
    err = (C - ref).abs().max().item()
    assert err <= 1e-13 * max(1.0, k**0.5) * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("m,n,k", [(128, 128, 8), (300, 200, 77), (1, 130, 1000), (1000, 1000, 512), (257, 511, 1030), (40, 18, 6)])
@pytest.mark.parametrize("alpha,beta", [(1.0, 0.0), (-1.0, 1.0), (0.5, -2.0)])
def test_gemm_nn_vs_torch(be, m, n, k, alpha, beta):
    """``lpgp_gemm_nn`` (B row-major k x n, swizzled 16-column boxes): ragged m / n / k, sub-block views with offsets."""
    g = torch.Generator(device="cuda").manual_seed(m + 2 * n + k)
    A = be.alloc_matrix(m, k).normal_(generator=g)
    Bfull = be.alloc_matrix(k + 3, n + 6).normal_(generator=g)
    B = Bfull[2 : 2 + k, 4 : 4 + n]  # a view at an even column offset inside a larger matrix (like L21 inside L)
    C = be.alloc_matrix(m, n).normal_(generator=g)
    ref = beta * C + alpha * (A @ B)
    be.gemm_nn(A, B, C, alpha, beta)
    err = (C - ref).abs().max().item()
    assert err <= 1e-13 * max(1.0, k**0.5) * max(1.0, ref.abs().max().item())


def test_gemm_nn_in_place_strip(be):
    """X <- X W in place (n <= 128: one CTA owns complete output rows), the leaf step of ``lpgp_trsm_rln``."""
    g = torch.Generator(device="cuda").manual_seed(5)
    for m, nb in ((7, 128), (1000, 128), (33000, 128), (300, 72)):
        X = be.alloc_matrix(m, 200).normal_(generator=g)
        W = be.alloc_matrix(128, 128).normal_(generator=g)
        ref = X[:, :nb] @ W[:nb, :nb]
        keep = X[:, nb:].clone()
        be.gemm_nn(X[:, :nb], W[:nb, :nb], X[:, :nb], 1.0, 0.0)
        assert (X[:, :nb] - ref).abs().max().item() <= 1e-12 * ref.abs().max().item()
        assert torch.equal(X[:, nb:], keep)


def test_gemm_nt_lower_only_touches_lower_tiles(be):
    n, k = 1000, 300
    A = be.alloc_matrix(n, k).normal_()
    C = be.alloc_matrix(n, n).zero_()
    be.gemm_nt(A, A, C, -1.0, 1.0, lower=True)
    ref = -(A @ A.T)
    low = torch.tril(torch.ones(n, n, dtype=torch.bool, device="cuda"))
    assert ((C - ref)[low]).abs().max().item() <= 1e-11
    # tiles strictly above the diagonal were skipped
    assert C[:128, 128:].abs().max().item() == 0.0


def _spd(n, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.randn(n, n + 16, dtype=torch.float64, device="cuda", generator=g)
    return X @ X.T / n + 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")


@pytest.mark.parametrize("n", [2, 64, 128, 130, 200, 384, 1000, 2304, 4098])
def test_potrf_vs_torch(be, n):
    G = _spd(n, n)
    f = be.DeviceFactor([n])
    f.L.copy_(G)
    f.potrf()
    L = torch.tril(f.L)
    L_ref = torch.linalg.cholesky(G)
    assert (L - L_ref).abs().max().item() <= 1e-12 * L_ref.abs().max().item()
    # valid square root + positive diagonal (probnum tests/test_linops/test_linop_decompositions.py:17-62)
    assert (L @ L.T - G).abs().max().item() <= 1e-12 * G.abs().max().item()
    assert (torch.diagonal(L) > 0).all()
    assert abs(f.logdet() - torch.logdet(G).item()) <= 1e-9 * max(1.0, abs(torch.logdet(G).item()))


@pytest.mark.parametrize("sizes", [(6000,), (2000, 3210), (130, 2, 5000)])
def test_potrf_lookahead_pipeline_matches_one_stream_recursion(be, sizes):
    """Ranges of >= 3 panels are factored by the two-stream right-looking pipeline with one panel of lookahead
    (cholesky.cu: potrf_lookahead); LPGP_OPT_NO_LOOKAHEAD selects the one-stream recursion.  Same factor up to
    rounding, for lpgp_potrf and for lpgp_chol_append (ragged segments)."""
    from linpde_gp_b200 import _lib

    n = sum(sizes)
    G = _spd(n, n)
    Ls = []
    for flag in (0, 1):
        assert _lib.lib.lpgp_set_option(_lib.OPT_NO_LOOKAHEAD, flag) == 0
        try:
            f, off = None, 0
            for s in sizes:
                f = be.DeviceFactor([s]) if f is None else f.extended(s)
                f.L[off : off + s, : off + s].copy_(G[off : off + s, : off + s])
                f.potrf() if off == 0 else f.append_last()
                off += s
            torch.cuda.synchronize()
            Ls.append(torch.tril(f.L).clone())
        finally:
            _lib.lib.lpgp_set_option(_lib.OPT_NO_LOOKAHEAD, 0)
    L_ref = torch.linalg.cholesky(G)
    sc = L_ref.abs().max().item()
    assert (Ls[0] - L_ref).abs().max().item() <= 1e-12 * sc
    assert (Ls[1] - L_ref).abs().max().item() <= 1e-12 * sc
    assert (Ls[0] - Ls[1]).abs().max().item() <= 1e-13 * sc


def _matern52_gram(n, ell, seed, noise=0.0):
    rng = np.random.default_rng(seed)
    x = np.sort(rng.uniform(0.0, 1.0, n))
    r = np.abs(x[:, None] - x[None, :]) * (np.sqrt(5.0) / ell)
    return (1.0 + r + r * r / 3.0) * np.exp(-r) + noise * np.eye(n)


@pytest.mark.parametrize("n,ell,noise,sizes", [(600, 0.2, 1e-12, (600,)), (600, 0.3, 1e-12, (256, 344)),
                                               (3000, 0.02, 1e-11, (3000,)), (2600, 0.03, 1e-11, (1000, 1600))])
def test_potrf_backward_error_ill_conditioned(be, n, ell, noise, sizes):
    """Backward stability of the blocked factorisation on nearly singular Gram matrices (cond 1e11 .. 1e13): with
    the residual-corrected panel solves (LPGP_OPT_TRSM_REFINE = 1, the default: gated per leaf by kappa_inf(L_kk) on the
    device; 3: every leaf) ``|L L^T - G| <= 1e-14 |G|`` like
    LAPACK's dpotrf -- the reference's factorisation, pn/linops/_linear_operator.py:860-865 -- whereas multiplying with
    the inverted diagonal blocks alone (option 0) leaves a residual of order cond(L_kk) eps."""
    from linpde_gp_b200 import _lib

    Gh = _matern52_gram(n, ell, seed=n, noise=noise)
    G = torch.as_tensor(Gh, device="cuda")
    sc = float(np.abs(Gh).max())
    errs = {}
    for mode in (1, 3, 0):
        assert _lib.lib.lpgp_set_option(_lib.OPT_TRSM_REFINE, mode) == 0
        try:
            f, off = None, 0
            for s in sizes:
                f = be.DeviceFactor([s]) if f is None else f.extended(s)
                f.L[off : off + s, : off + s].copy_(G[off : off + s, : off + s])
                f.potrf() if off == 0 else f.append_last()
                off += s
            L = torch.tril(f.L)
            errs[mode] = (L @ L.T - G).abs().max().item() / sc
        except np.linalg.LinAlgError:
            errs[mode] = float("inf")
        finally:
            _lib.lib.lpgp_set_option(_lib.OPT_TRSM_REFINE, 1)
    L_ref = torch.linalg.cholesky(G)
    err_ref = (L_ref @ L_ref.T - G).abs().max().item() / sc
    assert errs[1] <= max(1e-14, 8 * err_ref), (errs, err_ref)  # default: leaves with kappa_inf(L_kk) > 256 refined
    assert errs[3] <= max(1e-14, 8 * err_ref), (errs, err_ref)  # every leaf refined
    assert errs[1] <= errs[0]
    assert _lib.lib.lpgp_set_option(_lib.OPT_TRSM_REFINE, 4) != 0  # invalid value is rejected


def test_potrf_factors_nearly_singular_matrix_like_lapack(be):
    """Gram matrices with lambda_min = 1e-13 |G| (cond 1.4e15), which LAPACK factors: the refined blocked factorisation
    must factor them too (no spurious LinAlgError), with the same O(eps) backward error.  (A numpy emulation of the
    unrefined inverse-block scheme reports "not positive definite" for three of these four matrices.)"""
    for seed in range(4):
        Gh = _matern52_gram(600, 0.1, seed=seed, noise=1e-13)
        L_ref = np.linalg.cholesky(Gh)
        f = be.DeviceFactor([600])
        f.L.copy_(torch.as_tensor(Gh, device="cuda"))
        f.potrf()
        L = torch.tril(f.L).cpu().numpy()
        assert np.abs(L @ L.T - Gh).max() <= 8 * max(np.abs(L_ref @ L_ref.T - Gh).max(), 1e-15)


@pytest.mark.parametrize("refine", [False, True])
def test_trsm_rlt_refined_residual(be, refine):
    """lpgp_trsm_rlt_refined: residual |X L^T - B| at O(eps |X||L|) on an ill-conditioned factor."""
    import ctypes

    from linpde_gp_b200 import _lib

    n, m = 1500, 300
    Gh = _matern52_gram(n, 0.05, seed=1, noise=1e-11)
    f = be.DeviceFactor([n])
    f.L.copy_(torch.as_tensor(Gh, device="cuda"))
    f.potrf()
    L = torch.tril(f.L)
    B = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    X = be.alloc_matrix(m, n)
    X.copy_(B)
    st = f._struct()
    fn = _lib.lib.lpgp_trsm_rlt_refined if refine else _lib.lib.lpgp_trsm_rlt
    _lib.check(fn(ctypes.byref(st), n, ctypes.c_void_p(X.data_ptr()), m, X.stride(0), be._stream()), "trsm")
    resid = (X @ L.T - B).abs().max().item()
    bound = (X.abs() @ L.abs().T).abs().max().item()
    X_ref = torch.linalg.solve_triangular(L, B.T, upper=False).T
    resid_ref = (X_ref @ L.T - B).abs().max().item()
    if refine:
        assert resid <= max(64 * resid_ref, 1e-13 * bound), (resid, resid_ref, bound)
    else:
        assert resid <= 1e-6 * bound  # the fast path: cond(L_kk) eps, good enough for forward-error-bound uses


def test_potrf_lookahead_reports_first_failing_minor(be):
    n = 5000
    G = _spd(n, 3)
    G[3300, 3300] = -1.0
    f = be.DeviceFactor([n])
    f.L.copy_(G)
    with pytest.raises(np.linalg.LinAlgError) as ei:
        f.potrf()
    assert "3301" in str(ei.value)


def test_potrf_not_positive_definite_raises(be):
    n = 300
    G = _spd(n, 1)
    G[200, 200] = -1.0
    f = be.DeviceFactor([n])
    f.L.copy_(G)
    with pytest.raises(np.linalg.LinAlgError) as ei:
        f.potrf()
    assert "201" in str(ei.value)  # LAPACK-style info: order of the first non-PD leading minor


@pytest.mark.parametrize("n,m", [(130, 5), (1000, 300), (2304, 1), (2304, 700)])
def test_trsm_and_potrs_vs_torch(be, n, m):
    G = _spd(n, n + m)
    f = be.DeviceFactor([n])
    f.L.copy_(G)
    f.potrf()
    L_ref = torch.linalg.cholesky(G)
    B = torch.randn(m, n, dtype=torch.float64, device="cuda")
    X = be.alloc_matrix(m, n)
    X.copy_(B)
    f.trsm_rlt(X)
    ref = torch.linalg.solve_triangular(L_ref, B.T, upper=False).T
    assert (X - ref).abs().max().item() <= 1e-10 * ref.abs().max().item()
    Y = B[: min(m, 3)].clone()
    f.potrs(Y)
    ref2 = torch.cholesky_solve(B[: min(m, 3)].T.contiguous(), L_ref).T
    assert (Y - ref2).abs().max().item() <= 1e-9 * ref2.abs().max().item()


@pytest.mark.parametrize("sizes,m", [((130,), 5), ((200, 72, 300), 4), ((2, 4, 6, 130), 17), ((2304,), 700), ((5000, 2, 3002), 130)])
def test_trsm_rln_and_multi_rhs_potrs_vs_torch(be, sizes, m):
    """``lpgp_trsm_rln`` (X <- X L^{-1}) and ``lpgp_potrs`` with >= 4 right-hand sides (two blocked DMMA solves) against
    torch FP64, over ragged segments; the blocked path agrees with the single-vector substitution chains."""
    n = sum(sizes)
    G = _spd(n, n + m)
    f, off = None, 0
    for s in sizes:
        f = be.DeviceFactor([s]) if f is None else f.extended(s)
        f.L[off : off + s, : off + s].copy_(G[off : off + s, : off + s])
        f.potrf() if off == 0 else f.append_last()
        off += s
    L_ref = torch.linalg.cholesky(G)
    B = torch.randn(m, n, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(n + m))
    X = be.alloc_matrix(m, n)
    X.copy_(B)
    f.trsm_rln(X)
    ref = torch.linalg.solve_triangular(L_ref.T, B.T, upper=True).T
    assert (X - ref).abs().max().item() <= 1e-10 * ref.abs().max().item()
    Y = be.alloc_matrix(m, n)
    Y.copy_(B)
    f.potrs(Y)
    ref2 = torch.cholesky_solve(B.T.contiguous(), L_ref).T
    assert (Y - ref2).abs().max().item() <= 1e-9 * ref2.abs().max().item()
    assert (Y @ G - B).abs().max().item() <= 1e-11 * max(1.0, (Y.abs() @ G.abs()).max().item())
    y1 = f.potrs(B[:2].clone())  # < 4 rows: substitution chains
    assert (y1 - Y[:2]).abs().max().item() <= 1e-10 * Y[:2].abs().max().item()


@pytest.mark.parametrize("sizes", [(128, 128), (200, 72, 300), (2, 4, 6, 130), (1000, 24, 1500)])
def test_append_equals_full_factorisation(be, sizes):
    n = sum(sizes)
    G = _spd(n, n)
    f = None
    off = 0
    for s in sizes:
        f = be.DeviceFactor([s]) if f is None else f.extended(s)
        f.L[off : off + s, : off + s].copy_(G[off : off + s, : off + s])
        if off == 0:
            f.potrf()
        else:
            f.append_last()
        off += s
    L_ref = torch.linalg.cholesky(G)
    assert (torch.tril(f.L) - L_ref).abs().max().item() <= 1e-11 * L_ref.abs().max().item()
    b = torch.randn(1, n, dtype=torch.float64, device="cuda")
    x = f.potrs(b.clone())
    ref = torch.cholesky_solve(b.T.contiguous(), L_ref).T
    assert (x - ref).abs().max().item() <= 1e-9 * ref.abs().max().item()


def test_row_sumsq(be):
    A = be.alloc_matrix(37, 1001).normal_()
    out = be.row_sumsq(A, -1.0, 3.0)
    ref = 3.0 - (A * A).sum(dim=1)
    assert (out - ref).abs().max().item() <= 1e-11 * ref.abs().max().item()


@pytest.mark.parametrize("m,n", [(1, 1), (130, 77), (512, 4099), (1000, 1000), (7, 20000)])
def test_gemv_vs_torch(be, m, n):
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n)
    A = torch.randn(m, n + 3, dtype=torch.float64, device="cuda", generator=g)[:, :n]
    x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    z = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    y0 = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    y = be.gemv(A, x, y0.clone(), alpha=-0.75)
    assert (y - (y0 - 0.75 * (A @ x))).abs().max().item() <= 1e-12 * max(1.0, (A.abs() @ x.abs()).max().item())
    w0 = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    w = be.gemv(A, z, w0.clone(), alpha=1.5, trans=True)
    assert (w - (w0 + 1.5 * (A.T @ z))).abs().max().item() <= 1e-12 * max(1.0, (A.abs().T @ z.abs()).max().item())


@pytest.mark.parametrize("sizes", [(130,), (200, 72, 300), (2, 4, 6, 130), (12000,), (5000, 2, 7002)])
def test_potrs_residual_single_rhs(be, sizes):
    """Fused single-RHS substitution (one launch per leaf, in place): ``G x = b`` residual over ragged segments and
    sizes beyond the grid cap of the update CTAs (n > 9472)."""
    n = sum(sizes)
    G = _spd(n, n)
    f, off = None, 0
    for s in sizes:
        f = be.DeviceFactor([s]) if f is None else f.extended(s)
        f.L[off : off + s, : off + s].copy_(G[off : off + s, : off + s])
        f.potrf() if off == 0 else f.append_last()
        off += s
    b = torch.randn(2, n, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(n))
    x = f.potrs(b.clone())
    assert (x @ G - b).abs().max().item() <= 1e-11 * max(1.0, (x.abs() @ G.abs()).max().item())


@pytest.mark.parametrize("n", [2, 128, 300, 1664, 11000])
def test_trsv_forward_backward_vs_torch(be, n):
    import ctypes

    from linpde_gp_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(n)
    X = torch.randn(n, n + 8, dtype=torch.float64, device="cuda", generator=g)
    G = X @ X.T / n + 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
    f = be.DeviceFactor([n])
    f.L.copy_(G)
    f.potrf()
    L = torch.tril(f.L)
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    for trans in (0, 1):
        x = b.clone()
        st = f._struct()
        rc = _lib.lib.lpgp_trsv(ctypes.byref(st), trans, ctypes.c_void_p(x.data_ptr()), be._stream())
        assert rc == 0
        ref = torch.linalg.solve_triangular(L.T if trans else L, b[:, None], upper=bool(trans))[:, 0]
        assert (x - ref).abs().max().item() <= 1e-10 * ref.abs().max().item()
