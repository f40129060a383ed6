"""FP64 GEMM / triangular solve emulated on the INT8 tensor cores (csrc/ozaki.cu: tcgen05.mma.kind::i8, TMEM, TMA):
exactness of the digit splitting, the emulated product against torch FP64, the blocked solve against the DMMA path and
the posterior variance against the frozen reference output -- through the raw C ABI (ctypes)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from linpde_gp_b200 import backend

    return backend


def _rand(rng, m, n, spread=8.0):
    """entries with a wide dynamic range inside every row (the case the per-row scaling has to survive), a tenth of
    them (either sign) dozens of orders of magnitude below the row maximum, like far-apart kernel evaluations"""
    A = rng.standard_normal((m, n)) * np.exp2(rng.uniform(-spread, spread, (m, n)))
    tiny = rng.uniform(size=(m, n)) < 0.1
    A[tiny] *= 10.0 ** rng.uniform(-40, -10, size=int(tiny.sum()))
    return A


@pytest.mark.parametrize("S", [1, 3, 5, 7])
def test_split_is_exact_up_to_the_dropped_digits(be, S):
    rng = np.random.default_rng(S)
    m, n, kb = 200, 2048, 1024
    A = be.to_device(_rand(rng, m, n))
    A[3, :1024] = 0.0          # an all-zero K-block
    A[5, 7] = 2.0 ** 40        # one dominant entry
    A[6, :] = -A[6, :].abs()   # all-negative row
    A[7, 1::2] *= 1e-30        # entries (of both signs) far below the last digit kept
    A[8, :] = -1e-25
    A[8, 100] = 3.0
    P = be.OzakiPlanes(m, n, S, kb)
    P.split(A)
    R = P.reconstruct(slice(0, m), slice(0, n))
    # row maxima per K-block bound the truncation: |x - x_S| < 2^e 2^(-7 - 8 (S-1)) with max < 2^e <= 2 max
    mx = torch.stack([A[:, :1024].abs().amax(1), A[:, 1024:].abs().amax(1)], 1).repeat_interleave(1024, dim=1)
    err = (A - R).abs()
    assert torch.all(err <= 2.0 * mx * 2.0 ** (-7 - 8 * (S - 1)))
    assert torch.all(R <= A)   # truncation toward -inf
    if S == 7:  # 55 bits below 2^e: every entry of magnitude >= max / 4 (ulp >= 2^(e-55)) is reproduced exactly
        big = A.abs() >= 0.25 * mx
        assert torch.equal(A[big], R[big])
    assert torch.all(R[3, :1024] == 0.0)


@pytest.mark.parametrize("m,n,k,kb", [(128, 128, 1024, 1024), (100, 70, 2048, 1024), (300, 260, 4096, 2048), (128, 128, 512, 128)])
@pytest.mark.parametrize("S", [4, 6, 7])
def test_emulated_gemm_matches_fp64(be, m, n, k, kb, S):
    rng = np.random.default_rng(m + n + S)
    A, B = be.to_device(_rand(rng, m, k, 4.0)), be.to_device(_rand(rng, n, k, 4.0))
    C0 = be.to_device(rng.standard_normal((m, n)))
    PA, PB = be.OzakiPlanes(m, k, S, kb), be.OzakiPlanes(n, k, S, kb)
    PA.split(A)
    PB.split(B)
    C = be.alloc_matrix(m, n)
    C.copy_(C0)
    be.ozaki_gemm_nt(PA, PB, C, k, alpha=-1.5, beta=0.5)
    ref = 0.5 * C0 - 1.5 * (A @ B.T)
    # the product of the digits kept is exact: the error is the truncation, bounded against the per-K-block row maxima
    nkb = k // kb
    amax = A.abs().reshape(m, nkb, kb).amax(2)
    bmax = B.abs().reshape(n, nkb, kb).amax(2)
    bound = 1.5 * 4.0 * kb * (amax @ bmax.T) * (S + 1) * 2.0 ** (-7 - 8 * (S - 1)) + 1e-14 * ref.abs().max()
    assert torch.all((C - ref).abs() <= bound)
    if S == 7:  # digits of the planes multiplied in FP64 give the same matrix to rounding
        RA, RB = PA.reconstruct(slice(0, m), slice(0, k)), PB.reconstruct(slice(0, n), slice(0, k))
        ref7 = 0.5 * C0 - 1.5 * (RA @ RB.T)
        assert (C - ref7).abs().max() <= 1e-12 * (A.abs() @ B.abs().T).max()


@pytest.mark.parametrize("S", [1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("m,n", [(128, 128), (300, 260), (1000, 130)])
def test_level_orders_of_the_emulated_gemm_are_bit_identical(be, m, n, S):
    """LPGP_OPT_OZAKI_PAIR_LEVELS: two digit levels per pass over a K-block (operand tiles shared between the levels) or
    one -- the same exact int32 sums recombined in the same order, so the outputs must agree bit for bit; both against
    the FP64 product of the reconstructed digits.  (m = 128: no cluster; m > 128: clusters of two CTAs.)"""
    from linpde_gp_b200._lib import lib

    rng = np.random.default_rng(100 * S + m)
    k, kb = 3072, 1024
    A, B = be.to_device(_rand(rng, m, k, 4.0)), be.to_device(_rand(rng, n, k, 4.0))
    PA, PB = be.OzakiPlanes(m, k, S, kb), be.OzakiPlanes(n, k, S, kb)
    PA.split(A)
    PB.split(B)
    out = []
    try:
        assert lib.lpgp_set_option(7, 0) == 0  # the one-stream-per-CTA kernel (the CTA-pair kernel always pairs levels)
        for order in (0, 1):
            assert lib.lpgp_set_option(6, order) == 0
            C = be.alloc_matrix(m, n)
            C.fill_(0.25)
            be.ozaki_gemm_nt(PA, PB, C, k, alpha=1.0, beta=-2.0)
            torch.cuda.synchronize()
            out.append(C.clone())
    finally:
        assert lib.lpgp_set_option(6, 1) == 0
        assert lib.lpgp_set_option(7, 1) == 0
    assert torch.equal(out[0], out[1])
    if S == 7:
        RA, RB = PA.reconstruct(slice(0, m), slice(0, k)), PB.reconstruct(slice(0, n), slice(0, k))
        assert (out[1] - (RA @ RB.T - 0.5)).abs().max() <= 1e-12 * (A.abs() @ B.abs().T).max()


@pytest.mark.parametrize("S", [1, 2, 7])
@pytest.mark.parametrize("m,n", [(256, 128), (300, 260), (1000, 130)])
def test_cta_pair_kernel_is_bit_identical(be, m, n, S):
    """LPGP_OPT_OZAKI_CTA_PAIR: the tcgen05 cta_group::2 variant (M = 256 over the two CTAs of a cluster) against the
    one-stream-per-CTA kernel: same exact sums, same recombination order -> bit-identical."""
    from linpde_gp_b200._lib import lib

    rng = np.random.default_rng(1000 * S + m)
    k, kb = 3072, 1024
    A, B = be.to_device(_rand(rng, m, k, 4.0)), be.to_device(_rand(rng, n, k, 4.0))
    PA, PB = be.OzakiPlanes(m, k, S, kb), be.OzakiPlanes(n, k, S, kb)
    PA.split(A)
    PB.split(B)
    out = []
    try:
        for pair in (0, 1):
            assert lib.lpgp_set_option(7, pair) == 0
            C = be.alloc_matrix(m, n)
            C.fill_(0.25)
            be.ozaki_gemm_nt(PA, PB, C, k, alpha=1.0, beta=-2.0)
            torch.cuda.synchronize()
            out.append(C.clone())
    finally:
        assert lib.lpgp_set_option(7, 1) == 0
    if not torch.equal(out[0], out[1]):  # where: (128-row block, 64-column block) -> share of differing entries
        bad = (out[0] != out[1]).double()
        rb, cb = -(-m // 128), -(-n // 64)
        pad = torch.zeros((rb * 128, cb * 64), dtype=torch.float64, device=bad.device)
        pad[:m, :n] = bad
        share = pad.reshape(rb, 128, cb, 64).mean(dim=(1, 3)).cpu().numpy()
        raise AssertionError(f"CTA-pair kernel differs: max |d| = {(out[0] - out[1]).abs().max().item():.3e}, "
                             f"share of differing entries per (row block, column block):\n{np.round(share, 3)}")


def test_emulated_gemm_offsets_into_the_planes(be):
    rng = np.random.default_rng(0)
    kb, S = 1024, 6
    A, B = be.to_device(rng.standard_normal((384, 3072))), be.to_device(rng.standard_normal((512, 3072)))
    PA, PB = be.OzakiPlanes(384, 3072, S, kb), be.OzakiPlanes(512, 3072, S, kb)
    PA.split(A)
    PB.split(B)
    C = be.alloc_matrix(200, 130)
    be.ozaki_gemm_nt(PA, PB, C, 2048, rowA0=128, kA0=1024, rowB0=256, kB0=1024)
    ref = A[128:328, 1024:] @ B[256:386, 1024:].T
    assert (C - ref).abs().max() <= 1e-10 * ref.abs().max()


@pytest.mark.parametrize("n,m,S", [(2048, 300, 6), (3200, 257, 7), (4096, 1000, 5)])
def test_emulated_triangular_solve_matches_dmma_path(be, n, m, S):
    rng = np.random.default_rng(n)
    G = rng.standard_normal((n, n))
    G = G @ G.T / n + np.eye(n)
    f = be.DeviceFactor([n], reserve_rows=0)
    f.L.copy_(be.to_device(G))
    f.potrf()
    X0 = be.to_device(rng.standard_normal((m, n)))
    Xd, Xo = be.alloc_matrix(m, n), be.alloc_matrix(m, n)
    Xd.copy_(X0)
    Xo.copy_(X0)
    f.trsm_rlt(Xd)
    assert f.ozaki_eligible(1024)
    f.trsm_rlt_ozaki(Xo, S, 1024)
    tol = {5: 1e-8, 6: 1e-10, 7: 1e-12}[S]
    assert (Xo - Xd).abs().max() <= tol * Xd.abs().max()


def test_posterior_variance_with_the_emulated_solver_matches_the_frozen_reference(be):
    """Config 2 scaled to N = 4,096 (oracle/make_golden_large.py, outputs of the REAL reference): the variance computed
    with the INT8-emulated solve stays within the 1e-8 gate, and agrees with the DMMA path far below it."""
    import linpde_gp_b200 as lg
    from oracle import make_golden_large as mgl
    from tests import helpers

    path = os.path.join(os.path.dirname(__file__), "golden", "large_c2_4096.npz")
    if not os.path.exists(path):
        pytest.skip("large golden not generated")
    z = np.load(path)
    prob = mgl.large_problem(json.loads(bytes(z["problem_spec"]).decode()))
    k = helpers.api_kernel(prob["kernel"])
    prior = lg.GaussianProcess(lg.functions.Zero(input_shape=(2,)), k)
    batches = [(np.asarray(b["Y"]), np.asarray(b["X"]), helpers.api_op(b["L"])) for b in prob["blocks"]]
    post = lg.ConditionalGaussianProcess.from_observation_batches(prior, batches)
    Xt = np.asarray(prob["Xt"])
    prev = dict(be.VARIANCE_SOLVER)
    be.set_variance_solver(ozaki_slices=0)
    var_dmma = post.var(Xt)
    try:
        for S in (5, 6, 7):  # S = 5 (39 bits) is measurably outside the 1e-8 gate here (7e-8): never a default
            be.set_variance_solver(ozaki_slices=S)
            var_oz = post.var(Xt)
            assert np.max(np.abs(var_oz - z["var"])) <= {5: 1e-6, 6: 1e-8, 7: 1e-8}[S] * 4.0, S
            assert np.max(np.abs(var_oz - var_dmma)) <= {5: 1e-6, 6: 1e-9, 7: 1e-11}[S] * 4.0, S
    finally:
        be.set_variance_solver(ozaki_slices=prev["ozaki_slices"], kblock=prev["kblock"])


def test_non_finite_entries_propagate_like_an_fp64_gemm(be):
    """A NaN or an infinity in an operand turns the affected rows / columns of the product into NaN (an FP64 GEMM would
    give NaN or inf there); everything else stays exact."""
    rng = np.random.default_rng(9)
    m, n, k = 130, 140, 2048
    A, B = be.to_device(rng.standard_normal((m, k))), be.to_device(rng.standard_normal((n, k)))
    A[3, 1500] = float("nan")
    A[7, 10] = float("inf")
    B[5, 100] = -float("inf")
    PA, PB = be.OzakiPlanes(m, k, 7, 1024), be.OzakiPlanes(n, k, 7, 1024)
    PA.split(A)
    PB.split(B)
    C = be.alloc_matrix(m, n).zero_()
    be.ozaki_gemm_nt(PA, PB, C, k)
    bad = torch.zeros((m, n), dtype=torch.bool, device=C.device)
    bad[3, :] = bad[7, :] = True
    bad[:, 5] = True
    assert torch.all(torch.isnan(C[bad]))
    ref = torch.nan_to_num(A, nan=0.0, posinf=0.0, neginf=0.0) @ torch.nan_to_num(B, nan=0.0, posinf=0.0, neginf=0.0).T
    assert torch.all(torch.isfinite(C[~bad]))
    assert (C[~bad] - ref[~bad]).abs().max() <= 1e-12 * ref.abs().max()
